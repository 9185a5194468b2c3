"""Reverse-mode automatic differentiation over numpy ``longdouble`` arrays (x87 80-bit: 64-bit mantissa, eps 1.1e-19).

TEST INFRASTRUCTURE ONLY (see ``oracle/mfdgp_oracle.py``): the engine under ``oracle/mfdgp_truth.py``, the
extended-precision TRUTH that adjudicates between the fp64 oracle and the CUDA path wherever cond(K_zz) makes fp64
itself lose digits (the reference's default initialisation: cond 1e7 - 1e8).  With 11 more mantissa bits than fp64
its own rounding error is ~2000x below that of either contender, so ``|x - truth|`` measures x's error.

Nothing here is fast: matmul / Cholesky / triangular solves are numpy loops over longdouble (no BLAS / LAPACK exists
for it).  Sizes are kept to what finishes in seconds.
"""
import numpy as np

LD = np.longdouble
assert np.finfo(LD).eps < 2e-19, "numpy longdouble is not the 80-bit extended type on this platform"


class Var(object):
    """A node of the tape: value ``v`` (longdouble ndarray), parents and the function mapping the output gradient to
    the parents' gradients."""
    __slots__ = ("v", "g", "parents", "bw", "leaf")
    __array_ufunc__ = None      # ndarray (op) Var defers to Var.__r*__ instead of broadcasting over an object array

    def __init__(self, v, parents=(), bw=None, leaf=False):
        self.v = np.asarray(v, dtype=LD)
        self.g = None
        self.parents = parents
        self.bw = bw
        self.leaf = leaf

    # ---- operator sugar ----
    def __add__(self, o): return add(self, o)
    def __radd__(self, o): return add(o, self)
    def __sub__(self, o): return sub(self, o)
    def __rsub__(self, o): return sub(o, self)
    def __mul__(self, o): return mul(self, o)
    def __rmul__(self, o): return mul(o, self)
    def __truediv__(self, o): return div(self, o)
    def __rtruediv__(self, o): return div(o, self)
    def __neg__(self): return mul(self, -1.0)
    def __matmul__(self, o): return matmul(self, o)
    def __getitem__(self, idx): return getitem(self, idx)

    @property
    def shape(self): return self.v.shape

    @property
    def T(self): return transpose(self)


def leaf(x):
    """A differentiable input (from a torch tensor / ndarray / scalar)."""
    if hasattr(x, "detach"):
        x = x.detach().cpu().numpy()
    return Var(np.array(x, dtype=LD), leaf=True)


def const(x):
    if isinstance(x, Var):
        return x
    if hasattr(x, "detach"):
        x = x.detach().cpu().numpy()
    return Var(np.asarray(x, dtype=LD))


def _unbroadcast(g, shape):
    """Sum the gradient of a broadcast operand back to the operand's shape."""
    g = np.asarray(g, dtype=LD)
    while g.ndim > len(shape):
        g = g.sum(axis=0)
    for ax, n in enumerate(shape):
        if n == 1 and g.shape[ax] != 1:
            g = g.sum(axis=ax, keepdims=True)
    return g.reshape(shape)


def add(a, b):
    a, b = const(a), const(b)
    return Var(a.v + b.v, (a, b), lambda g: (_unbroadcast(g, a.v.shape), _unbroadcast(g, b.v.shape)))


def sub(a, b):
    a, b = const(a), const(b)
    return Var(a.v - b.v, (a, b), lambda g: (_unbroadcast(g, a.v.shape), _unbroadcast(-g, b.v.shape)))


def mul(a, b):
    a, b = const(a), const(b)
    return Var(a.v * b.v, (a, b), lambda g: (_unbroadcast(g * b.v, a.v.shape), _unbroadcast(g * a.v, b.v.shape)))


def div(a, b):
    a, b = const(a), const(b)
    out = a.v / b.v
    return Var(out, (a, b), lambda g: (_unbroadcast(g / b.v, a.v.shape), _unbroadcast(-g * out / b.v, b.v.shape)))


def exp(a):
    out = np.exp(a.v)
    return Var(out, (a,), lambda g: (g * out,))


def log(a):
    return Var(np.log(a.v), (a,), lambda g: (g / a.v,))


def sqrt(a):
    out = np.sqrt(a.v)
    return Var(out, (a,), lambda g: (g / (2 * out),))


def square(a):
    return Var(a.v * a.v, (a,), lambda g: (2 * g * a.v,))


def softplus(a):
    """log(1 + e^x) and its derivative sigmoid(x), overflow-safe."""
    x = a.v
    out = np.where(x > 0, x + np.log1p(np.exp(-np.abs(x))), np.log1p(np.exp(-np.abs(x))))
    sig = np.where(x >= 0, 1 / (1 + np.exp(-np.abs(x))), np.exp(-np.abs(x)) / (1 + np.exp(-np.abs(x))))
    return Var(out, (a,), lambda g: (g * sig,))


def sigmoid(a):
    x = a.v
    e = np.exp(-np.abs(x))
    out = np.where(x >= 0, 1 / (1 + e), e / (1 + e))
    return Var(out, (a,), lambda g: (g * out * (1 - out),))


def clamp_min(a, c):
    """torch.clamp(min=c): value max(x, c), gradient passes where x >= c (torch's convention: inclusive)."""
    c = LD(c)
    m = a.v >= c
    return Var(np.where(m, a.v, c), (a,), lambda g: (g * m,))


def sum_(a, axis=None, keepdims=False):
    out = a.v.sum(axis=axis, keepdims=keepdims)

    def bw(g):
        g = np.asarray(g, dtype=LD)
        if axis is not None and not keepdims:
            g = np.expand_dims(g, axis)
        return (np.broadcast_to(g, a.v.shape).copy(),)
    return Var(out, (a,), bw)


def transpose(a):
    return Var(a.v.T, (a,), lambda g: (g.T,))


def reshape(a, shape):
    return Var(a.v.reshape(shape), (a,), lambda g: (g.reshape(a.v.shape),))


def getitem(a, idx):
    def bw(g):
        out = np.zeros(a.v.shape, dtype=LD)
        np.add.at(out, idx, g)
        return (out,)
    return Var(a.v[idx], (a,), bw)


def repeat_interleave(a, n, axis=0):
    idx = np.repeat(np.arange(a.v.shape[axis]), n)
    return getitem(a, idx) if axis == 0 else getitem(a, (slice(None),) * axis + (idx,))


def cat(parts, axis=0):
    parts = [const(p) for p in parts]
    sizes = [p.v.shape[axis] for p in parts]

    def bw(g):
        out, o = [], 0
        for n in sizes:
            sl = [slice(None)] * g.ndim
            sl[axis] = slice(o, o + n)
            out.append(g[tuple(sl)])
            o += n
        return tuple(out)
    return Var(np.concatenate([p.v for p in parts], axis=axis), tuple(parts), bw)


def tril(a):
    return Var(np.tril(a.v), (a,), lambda g: (np.tril(g),))


def matmul(a, b):
    a, b = const(a), const(b)
    return Var(a.v @ b.v, (a, b), lambda g: (g @ b.v.T if b.v.ndim == 2 else np.outer(g, b.v),
                                             a.v.T @ g if a.v.ndim == 2 else np.outer(a.v, g)))


# ---- dense factorisation and triangular solves in longdouble (plain loops: no LAPACK for this type) ----
def _chol(A):
    n = A.shape[0]
    L = np.zeros_like(A)
    for j in range(n):
        d = A[j, j] - L[j, :j] @ L[j, :j]
        if not d > 0:
            raise RuntimeError("NotPSDError (longdouble truth): pivot %d = %r" % (j, d))
        L[j, j] = np.sqrt(d)
        if j + 1 < n:
            L[j + 1:, j] = (A[j + 1:, j] - L[j + 1:, :j] @ L[j, :j]) / L[j, j]
    return L


def _solve_lower(L, B):
    """L^-1 B by forward substitution (B: n x r or n)."""
    X = np.array(B, dtype=LD, copy=True)
    for i in range(L.shape[0]):
        if i:
            X[i] -= L[i, :i] @ X[:i]
        X[i] /= L[i, i]
    return X


def _solve_lower_t(L, B):
    """L^-T B by backward substitution."""
    X = np.array(B, dtype=LD, copy=True)
    n = L.shape[0]
    for i in range(n - 1, -1, -1):
        if i + 1 < n:
            X[i] -= L[i + 1:, i] @ X[i + 1:]
        X[i] /= L[i, i]
    return X


def cholesky(a):
    """Lower Cholesky factor of a symmetric matrix; backward after Murray (2016):
    Abar = sym(L^-T Phi(L^T Lbar) L^-1), Phi = lower triangle with the diagonal halved."""
    L = _chol(a.v)

    def bw(g):
        P = np.tril(L.T @ np.tril(g))
        P[np.diag_indices_from(P)] *= LD(0.5)
        Y = _solve_lower_t(L, P)                  # L^-T Phi
        S = _solve_lower_t(L, Y.T).T              # (L^-T Phi) L^-1
        return ((S + S.T) / 2,)
    return Var(L, (a,), bw)


def solve_lower(L, b):
    """X = L^-1 B.  Bbar = L^-T Xbar,  Lbar = -tril(Bbar X^T)."""
    L, b = const(L), const(b)
    X = _solve_lower(L.v, b.v)

    def bw(g):
        Bb = _solve_lower_t(L.v, g)
        Lb = -np.tril(Bb @ X.T) if X.ndim == 2 else -np.tril(np.outer(Bb, X))
        return (Lb, Bb)
    return Var(X, (L, b), bw)


def solve_lower_t(L, b):
    """X = L^-T B.  Bbar = L^-1 Xbar,  Lbar = -tril(X Bbar^T)."""
    L, b = const(L), const(b)
    X = _solve_lower_t(L.v, b.v)

    def bw(g):
        Bb = _solve_lower(L.v, g)
        Lb = -np.tril(X @ Bb.T) if X.ndim == 2 else -np.tril(np.outer(X, Bb))
        return (Lb, Bb)
    return Var(X, (L, b), bw)


def backward(out):
    """d out / d every node reachable from the scalar ``out`` (gradients land in ``.g`` of the leaves)."""
    order, seen = [], set()
    stack = [(out, False)]
    while stack:
        node, done = stack.pop()
        if done:
            order.append(node)
            continue
        if id(node) in seen:
            continue
        seen.add(id(node))
        stack.append((node, True))
        for p in node.parents:
            if id(p) not in seen:
                stack.append((p, False))
    for n in order:
        n.g = None
    out.g = np.ones_like(out.v)
    for node in reversed(order):
        if node.bw is None or node.g is None:
            continue
        for p, gp in zip(node.parents, node.bw(node.g)):
            if gp is None:
                continue
            gp = np.asarray(gp, dtype=LD)
            p.g = gp if p.g is None else p.g + gp
        if not node.leaf:
            node.g = None if node is not out else node.g
