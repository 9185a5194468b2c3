"""Exception types with the names upstream raises (linear_operator.utils.errors), so that ``except NotPSDError`` code
written against the reference keeps working."""


class NotPSDError(RuntimeError):
    pass


class NanError(RuntimeError):
    pass
