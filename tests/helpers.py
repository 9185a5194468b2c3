"""Shared builders for the parity tests: random but well-defined MFDGP states in GPyTorch's state_dict naming."""
import numpy as np
import torch

from oracle import mfdgp_oracle as O


def random_state(M, d, L, seed=0, ls=0.5, lq_scale=0.05, dtype=torch.float64):
    """A random L-layer MFDGP state with shared inducing inputs (so quirk Q4 holds: Z_l = [Z, m_{l-1}])."""
    g = torch.Generator().manual_seed(seed)
    Zx = torch.rand(M, d, generator=g, dtype=dtype)
    sd = {}
    for l in range(L):
        p = "hidden_layer_%d." % l
        c = p + "covar_module."

        def rawpos(lo, hi, shape):
            v = lo + (hi - lo) * torch.rand(shape, generator=g, dtype=dtype)
            return O.inv_softplus(v)

        if l == 0:
            sd[c + "raw_outputscale"] = rawpos(0.8, 1.5, ())
            sd[c + "base_kernel.raw_lengthscale"] = rawpos(0.7 * ls, 1.3 * ls, (1, d))
            Z = Zx.clone()
        else:
            sd[c + "kernels.0.kernels.0.raw_outputscale"] = rawpos(0.7, 1.3, ())
            sd[c + "kernels.0.kernels.0.base_kernel.raw_lengthscale"] = rawpos(1.5 * ls, 3.0 * ls, (1, d))
            sd[c + "kernels.0.kernels.1.kernels.0.raw_variance"] = rawpos(0.5, 1.2, (1, 1))
            sd[c + "kernels.0.kernels.1.kernels.1.raw_outputscale"] = rawpos(0.6, 1.2, ())
            sd[c + "kernels.0.kernels.1.kernels.1.base_kernel.raw_lengthscale"] = rawpos(0.7, 1.3, (1, 1))
            sd[c + "kernels.1.raw_outputscale"] = rawpos(0.05, 0.2, ())
            sd[c + "kernels.1.base_kernel.raw_lengthscale"] = rawpos(0.7 * ls, 1.3 * ls, (1, d))
            Z = torch.cat([Zx, torch.zeros(M, 1, dtype=dtype)], 1)    # stored last column is ignored at run time
        sd[p + "variational_strategy.inducing_points"] = Z
        q = p + "variational_strategy._variational_distribution."
        sd[q + "variational_mean"] = torch.randn(M, generator=g, dtype=dtype)
        Lq = torch.tril(torch.randn(M, M, generator=g, dtype=dtype)) * lq_scale / np.sqrt(M)
        Lq = Lq + torch.diag(0.05 + 0.1 * torch.rand(M, generator=g, dtype=dtype))
        # strictly-upper garbage must be ignored (CholeskyVariationalDistribution masks with tril)
        sd[q + "chol_variational_covar"] = Lq + torch.triu(torch.randn(M, M, generator=g, dtype=dtype), 1)
        sd["hidden_layer_likelihood_%d.noise_covar.raw_noise" % l] = torch.randn(1, generator=g, dtype=dtype)
    noise_upper = [0.1 + 0.02 * l for l in range(L)]
    return sd, noise_upper


def param_keys(sd):
    return [k for k, v in sd.items() if v.dtype.is_floating_point and "inducing_points" not in k]


def clone_state(sd, device=None, requires_grad=False):
    out = {}
    for k, v in sd.items():
        t = v.detach().clone()
        if device is not None:
            t = t.to(device)
        if requires_grad and t.dtype.is_floating_point and "inducing_points" not in k:
            t.requires_grad_(True)
        out[k] = t
    return out


def relerr(a, b):
    a = a.detach().cpu().double()
    b = b.detach().cpu().double()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-300))


# ------------------------------------------------------------------------------------------------------------
# fixtures from the reference's examples
# ------------------------------------------------------------------------------------------------------------
def forrester_data():
    """Deterministic data of examples/example_acquisition_mfdgp_forrester/...py:51-104 (config C2)."""
    def mf1(x):
        return ((6 * x - 2) ** 2) * np.sin(12 * x - 4)

    def mf0(x):
        return 0.5 * mf1(x) + 10 * (x - 0.5) + 5
    x0 = np.linspace(0, 1.0, 12).reshape(12, 1)
    x1 = np.array([0.1, 0.3, 0.5, 0.7]).reshape(4, 1)
    out = {}
    for name, f0, f1 in (("obj1", mf0, mf1), ("obj2", lambda x: -mf0(x), lambda x: -mf1(x)),
                         ("con1", lambda x: np.sin(x * np.pi * 2.5), lambda x: np.cos(x * np.pi * 2.5))):
        y0, y1 = f0(x0), f1(x1)
        mean, std = np.mean(np.vstack((y1, y0))), np.std(np.vstack((y1, y0)))
        out[name] = torch.cat((torch.from_numpy((y1 - mean) / std), torch.from_numpy((y0 - mean) / std)), 0).double()
    x = torch.cat((torch.from_numpy(x1), torch.from_numpy(x0)), 0).double()
    fid = torch.cat((torch.ones(4).double(), torch.zeros(12).double()))[:, None]
    return x, out, fid


def synthetic_data(n_per_fid, d, seed=0):
    """Seeded closed-form multi-fidelity data (SURVEY.md §8d C4 recipe), ordering: all fidelities concatenated."""
    g = torch.Generator().manual_seed(seed)
    L = len(n_per_fid)
    xs, ys, fs = [], [], []
    for l, n in enumerate(n_per_fid):
        x = torch.rand(n, d, generator=g, dtype=torch.float64)
        y0 = torch.sin(2 * np.pi * x).sum(1) / np.sqrt(d)
        y1 = 0.8 * y0 + 0.2 * torch.cos(np.pi * x.sum(1))
        y2 = y1 ** 2 - 0.5 * y1 + 0.1 * x[:, 0]
        y = [y0, y1, y2][min(l, 2)] + 1e-2 * torch.randn(n, generator=g, dtype=torch.float64)
        xs.append(x); ys.append(y[:, None]); fs.append(torch.full((n, 1), float(l), dtype=torch.float64))
    del L
    return torch.cat(xs), torch.cat(ys), torch.cat(fs)


def oracle_view(model):
    """CPU copy of a product model's parameters for the oracle: (sd, noise_lower, noise_upper, samples)."""
    sd = {k: v.detach().cpu().clone() for k, v in model.state_dict().items() if "previous_layer" not in k}
    L = model.num_fidelities
    lo = [float(sd["hidden_layer_likelihood_%d.noise_covar.raw_noise_constraint.lower_bound" % l]) for l in range(L)]
    up = [float(sd["hidden_layer_likelihood_%d.noise_covar.raw_noise_constraint.upper_bound" % l]) for l in range(L)]
    samples = [getattr(model, "hidden_layer_%d" % l).samples.detach().cpu() for l in range(L)]
    return sd, lo, up, samples


def parity_tol(model, base=1e-10):
    """fp64 parity bar: 1e-10 relative on well-conditioned inputs, cond(K_zz + jitter I) * eps-scaled otherwise
    (SURVEY.md §7 'Hard parts': two algebraically identical fp64 formulations already differ by ~cond * eps)."""
    sd, _, _, _ = oracle_view(model)
    worst = 1.0
    for l in range(model.num_fidelities):
        Z = O.layer_inducing_points(sd, l)
        P = O.layer_kernel(sd, l, Z, Z) + O.JITTER * torch.eye(Z.shape[0], dtype=torch.float64)
        worst = max(worst, float(torch.linalg.cond(P)))
    return max(base, 20 * 2.2e-16 * worst), worst


def model_from_state(sd, noise_upper, L, samples=None, device="cuda:0", noise_lower=None):
    """A product MFDGP whose parameters, inducing inputs, noise bounds (and eval samples) equal the given oracle
    state: the bridge from golden vectors / random states to the drop-in classes."""
    from mobocmf_b200.models.mfdgp import MFDGP
    Zx = sd["hidden_layer_0.variational_strategy.inducing_points"]
    M, d = Zx.shape
    fid = (torch.arange(M) % L).double()[:, None]
    y = torch.linspace(-1.0, 1.0, M, dtype=torch.float64)[:, None]
    model = MFDGP(Zx.clone(), y, fid, L, init_lengthscale=0.5)
    model.double()
    own = model.state_dict()
    with torch.no_grad():
        for k, v in sd.items():
            if k in own and "inducing_points" not in k:
                own[k].copy_(v.reshape(own[k].shape))
        for l in range(L):
            c = getattr(model, "hidden_layer_likelihood_%d" % l).noise_covar.raw_noise_constraint
            c.upper_bound.copy_(torch.as_tensor(noise_upper[l], dtype=torch.float64))
            c.lower_bound.copy_(torch.as_tensor(O.NOISE_LOWER if noise_lower is None else noise_lower[l],
                                                dtype=torch.float64))
            if samples is not None:
                layer = getattr(model, "hidden_layer_%d" % l)
                layer.samples = samples[l].clone()
                layer.num_samples_for_acquisition = samples[l].shape[0]
                model.num_samples_for_acquisition = samples[l].shape[0]
    return model.to(device)


# ------------------------------------------------------------------------------------------------------------
# extended-precision adjudication (oracle/mfdgp_truth.py): the bar for ill-conditioned cases
# ------------------------------------------------------------------------------------------------------------
ADJ_C = 10.0        # the CUDA path may be at most this many times further from the exact answer than the fp64 oracle
ADJ_FLOOR = 1e-10   # ... or within the north-star tolerance of it
ADJ_KR = 64.0       # ... or within the effect of rounding every covariance entry by ADJ_KR * eps, whichever is largest
EPS64 = 2.220446049250313e-16


def adjudicate(tag, cuda, oracle, truth, sens=None, c=ADJ_C, floor=ADJ_FLOOR, kr=ADJ_KR, report=None):
    """Asserts |cuda - truth| <= max(floor, c * |oracle - truth|, kr * sens) (max-norm, relative to max |truth|).
    ``truth`` is a numpy longdouble array from oracle/mfdgp_truth.py, whose own rounding error is ~2000x below fp64's;
    ``sens`` is how far the truth of THIS quantity moves when every covariance entry is perturbed by a relative eps64
    (oracle/mfdgp_truth.covariance_rounding).

    Why three terms.  c * |oracle - truth| is the comparison that matters: the CUDA path must be about as close to the
    exact answer as the reference-shaped fp64 program.  But the oracle's REALISED error is not a bound: its triangular
    solves are far more accurate than their worst case (it lands within 1e-12 of the truth at cond 1e7 in some cases),
    while the CUDA path multiplies by explicit triangular inverses (what puts the work on the DMMA pipe), whose error
    sits near the a-priori bound of a solve.  kr * sens is that bound made specific to the quantity: the forward error
    of a program whose result is exact for covariance matrices rounded by kr * eps (the backward error of a Cholesky
    solve of this size).  It replaces a flat eps * cond(K_zz) allowance, which is blind to cancellation: the gradient
    of layer 0's outputscale on the Forrester fixture is 354 while its two shares (through K_zz, through K_zx) are
    ~1e6 each, the CUDA path gets both to 0.03 eps cond, the sum to 87 eps cond, and rounding K alone already moves it
    by 11 eps cond (tools/parity_diag.py, profiles/r02_parity_adjudication.txt).  The bars this replaced were 1e-2 /
    1e-3 relative."""
    from oracle import mfdgp_truth as T
    ec, eo = T.err_vs(cuda, truth), T.err_vs(oracle, truth)
    if report is not None:
        report.append((tag, ec, eo, sens))
    bar = max(floor, c * eo, (kr * sens) if sens else 0.0)
    assert ec <= bar, "%s: |cuda - truth| = %.2e but |oracle - truth| = %.2e (covariance-rounding sensitivity %s, bar %.2e)" % (
        tag, ec, eo, "%.1e" % sens if sens else "-", bar)
    return ec, eo


def state_cond(sd, L):
    """max over the layers of cond(K(Z_l, Z_l) + jitter I) for an oracle-format state."""
    worst = 1.0
    for l in range(L):
        Z = O.layer_inducing_points(sd, l)
        P = O.layer_kernel(sd, l, Z, Z) + O.JITTER * torch.eye(Z.shape[0], dtype=torch.float64)
        worst = max(worst, float(torch.linalg.cond(P)))
    return worst


def adjudicate_state(sd, lo, up, loss_cuda, grads_cuda, loss_oracle, grads_oracle, L, xb, yb, fb, eps, num_data, S,
                     c=ADJ_C, floor=ADJ_FLOOR, verbose=True, only_hf=False):
    """Loss (= -ELBO) and EVERY gradient of one ELBO step on the oracle-format state ``sd``: CUDA vs fp64 oracle,
    adjudicated by the longdouble truth evaluated on the same parameters, minibatch and normals.
    grads_*: {parameter name: tensor}.  Returns [(tag, err_cuda, err_oracle, sensitivity)]."""
    import numpy as np
    from oracle import mfdgp_truth as T
    names = sorted(grads_oracle)
    sd = {k: v.detach() for k, v in sd.items()}
    with torch.no_grad():
        cond = state_cond(sd, L)
    loss_t, _, grads_t = T.elbo_step_truth(sd, names, L, up, xb, yb, fb, eps, num_data, S, noise_lower=lo,
                                           only_hf=only_hf)
    sens = T.elbo_step_rounding_sensitivity(sd, names, L, up, xb, yb, fb, eps, num_data, S, (loss_t, grads_t),
                                            noise_lower=lo, only_hf=only_hf)
    rep = []
    adjudicate("loss", loss_cuda, loss_oracle, loss_t, sens["loss"], c, floor, report=rep)
    for n in names:
        gc, go, gt = grads_cuda[n], grads_oracle[n], grads_t[n]
        if "chol_variational_covar" in n:
            gc, go, gt = torch.tril(gc), torch.tril(go), np.tril(gt)
        adjudicate(n, gc, go, gt, sens[n], c, floor, report=rep)
    if verbose:
        worst = max(rep, key=lambda r: r[1])
        above = [r for r in rep if r[1] > max(floor, c * r[2])]        # decided by the sensitivity term
        print("adjudicated %d quantities at cond %.1e: worst |cuda - truth| %.2e = %.2f eps cond (%s), there "
              "|oracle - truth| %.2e and covariance-rounding sensitivity %.2e; %d quantities beyond %g x oracle: "
              "max |cuda - truth| / sensitivity there %.1f"
              % (len(rep), cond, worst[1], worst[1] / (EPS64 * cond), worst[0], worst[2], worst[3], len(above), c,
                 max([r[1] / r[3] for r in above] or [0.0])))
    return rep


def adjudicate_step(model, loss_cuda, grads_cuda, loss_oracle, grads_oracle, L, xb, yb, fb, eps, num_data, S, **kw):
    """``adjudicate_state`` on a product model's parameters."""
    sd, lo, up, _ = oracle_view(model)
    return adjudicate_state(sd, lo, up, loss_cuda, grads_cuda, loss_oracle, grads_oracle, L, xb, yb, fb, eps,
                            num_data, S, **kw)
