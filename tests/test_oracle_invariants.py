"""CPU: the oracle against the invariants the reference offers without GPyTorch and against itself."""
import torch

from oracle import mfdgp_oracle as O
from tests.helpers import random_state, forrester_data


def test_acquisition_fidelity0_equals_single_prediction():
    # S identical copies at fidelity 0  =>  predict_for_acquisition == single prediction
    sd, up = random_state(12, 2, 2, seed=1, ls=0.3)
    samples = [torch.randn(25, 1), torch.randn(25, 1)]
    X = torch.rand(9, 2, dtype=torch.float64)
    mu, var = O.predict_for_acquisition(sd, 2, up, samples, X, 0)
    mu1, var1 = O.predict(sd, 2, up, X, 0, training=False)
    assert torch.allclose(mu, mu1.reshape(-1), rtol=1e-12, atol=1e-14)
    assert torch.allclose(var, var1.reshape(-1), rtol=1e-9, atol=1e-13)


def test_cond_equals_uncond_gives_zero_acquisition():
    sd, up = random_state(12, 2, 2, seed=2, ls=0.3)
    samples = [torch.randn(25, 1), torch.randn(25, 1)]
    mod = dict(sd=sd, num_layers=2, noise_upper=up, samples=samples)
    X = torch.rand(9, 1, 2, dtype=torch.float64)
    assert float(O.jes_mfdgp(mod, mod, X, 1).abs().max()) == 0.0


def test_shortcut_and_kl_zero():
    sd, up = random_state(10, 1, 2, seed=3, ls=0.2)
    Z = sd["hidden_layer_0.variational_strategy.inducing_points"]
    m, Lq = O.variational_q(sd, 0)
    mean, var = O.layer_q(sd, 0, Z)                       # quirk Q4
    assert torch.equal(mean, m) and torch.allclose(var, (Lq @ Lq.T).diagonal())
    Z1 = O.layer_inducing_points(sd, 1)
    assert torch.equal(Z1[:, -1], m)                      # Z_1's last column is m_0 bit-for-bit
    # KL = 0 when q = prior
    Lp = O.prior_cholesky(sd, 0)
    q = "hidden_layer_0.variational_strategy._variational_distribution."
    sd[q + "variational_mean"] = torch.zeros_like(m)
    sd[q + "chol_variational_covar"] = Lp
    assert abs(float(O.kl_layer(sd, 0))) < 1e-9


def test_literal_eval_covariance_diag_matches_diag_only_branch():
    sd, up = random_state(14, 2, 2, seed=4, ls=0.3)
    X = torch.rand(11, 3, dtype=torch.float64)
    a = O.layer_q(sd, 1, X, training=False)
    b = O.layer_q(sd, 1, X, training=False, literal_eval_cov=True)
    assert torch.allclose(a[1], b[1], rtol=1e-9, atol=1e-13)


def test_tiled_multisample_equals_mean_of_single_sample_elbos():
    sd, up = random_state(12, 2, 3, seed=5, ls=0.3)
    g = torch.Generator().manual_seed(1)
    B, S = 7, 4
    x = torch.rand(B, 2, generator=g, dtype=torch.float64)
    y = torch.randn(B, 1, generator=g, dtype=torch.float64)
    fid = torch.randint(0, 3, (B, 1), generator=g).double()
    eps_t = [None] + [torch.randn(B * S, generator=g, dtype=torch.float64) for _ in range(2)]
    eps_list = [[None] + [eps_t[l].reshape(B, S)[:, s].reshape(1, B) for l in (1, 2)] for s in range(S)]
    a, kla = O.elbo_step_loss_tiled(sd, 3, up, x, y, fid, eps_t, 50, S)
    b, klb = O.elbo_step_loss_multisample(sd, 3, up, x, y, fid, eps_list, 50)
    assert abs(float(a - b)) < 1e-9 * abs(float(b)) and abs(float(kla - klb)) < 1e-12 * abs(float(klb))


def test_forrester_elbo_is_finite_and_differentiable():
    x, ys, fid = forrester_data()
    torch.manual_seed(0)
    sd, lo, up, samples = O.init_state_dict(x, ys["obj1"], fid, 2)
    for k in sd:
        if sd[k].dtype.is_floating_point and "inducing_points" not in k:
            sd[k].requires_grad_(True)
    loss, kl = O.elbo_step_loss(sd, 2, up, x, ys["obj1"], fid, [None, torch.randn(1, 16)], 16, noise_lower=lo)
    loss.backward()
    assert torch.isfinite(loss) and all(torch.isfinite(v.grad).all() for v in sd.values() if v.requires_grad)
