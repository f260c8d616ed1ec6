"""CPU checks of the arithmetic the tensor-core LSTM path uses on the GPU
(csrc/kernels_lstm_tc.cu, csrc/tc_core.cuh), emulated in numpy float32:

  * the fp16 (hi, lo) operand split and the three-product evaluation of an f32 product;
  * the shared-reciprocal LSTM cell (5 exponentials + 2 reciprocals per unit);
  * the degree-5 polynomial 2^x of the FMA-pipe fallback (PB_TC_NPOLY);
  * the geometry of the (scale, shift) uncertainty triangle of the scaler guard.

These are accuracy statements the GPU kernels rely on, not parity tests: the
parity of the GPU path is established against the exact kernels / the oracle in the -m gpu suite.
"""
import numpy as np

f32 = np.float32


def _split(v):
    """hi = value truncated to 11 significant bits, lo = fp16(value - hi)  (split_pair)."""
    v = np.asarray(v, f32)
    hi = (v.view(np.uint32) & np.uint32(0xFFFFE000)).view(f32)
    lo = (v - hi).astype(np.float16).astype(f32)
    return hi.astype(np.float16).astype(f32), lo


def test_split_fp16_three_products_reach_f32_accuracy():
    rng = np.random.default_rng(0)
    a = rng.uniform(-1, 1, (256, 160)).astype(f32)            # hidden states
    w = rng.normal(0, 0.6, (160, 256)).astype(f32)            # weights
    ah, al = _split(a)
    wh, wl = _split(w)
    exact = a.astype(np.float64) @ w.astype(np.float64)
    three = ah.astype(np.float64) @ wh + ah.astype(np.float64) @ wl + al.astype(np.float64) @ wh
    one = ah.astype(np.float64) @ wh
    err3 = np.abs(three - exact).max()
    err1 = np.abs(one - exact).max()
    chain = np.zeros((256, 256), f32)                          # the exact kernels' f32 fma chain
    for k in range(160):
        chain = (chain.astype(np.float64) + a[:, k:k + 1].astype(np.float64) * w[k].astype(np.float64)).astype(f32)
    err_chain = np.abs(chain - exact).max()
    assert err3 < 4e-6                   # same size as the f32 chain's own rounding noise
    assert err3 < 4 * err_chain + 1e-6
    assert err1 > 100 * err3             # the coarse probe perturbs by ~2^-11


def _cell_shared_rcp(zi, zf, zc, zo, c, rng):
    L = f32(1.4426950408889634)

    def ex2(x):                          # MUFU.EX2: relative error 2^-22, flush below 2^-126
        y = np.exp2(x.astype(np.float64)) * (1 + rng.uniform(-1, 1, x.shape) * 2.0 ** -22)
        y = y.astype(f32)
        y[x < -126] = 0
        return y

    def rcp(x):                          # MUFU.RCP: 1 ulp
        return (1 / x.astype(np.float64) * (1 + rng.uniform(-1, 1, x.shape) * 2.0 ** -23)).astype(f32)

    ei = ex2(np.minimum(-zi * L, f32(30)))
    ef = ex2(np.minimum(-zf * L, f32(30)))
    eg = ex2(np.minimum(f32(-2) * zc * L, f32(30)))
    eo = ex2(np.minimum(-zo * L, f32(30)))
    af = f32(1) + ef
    p = (f32(1) + ei) * (f32(1) + eg)
    r = rcp(p * af)
    ig = (f32(1) - eg) * af * r
    cn = ((p * r).astype(np.float64) * c + ig).astype(f32)
    ec = ex2(np.minimum(f32(-2) * cn * L, f32(30)))
    r2 = rcp((f32(1) + eo) * (f32(1) + ec))
    return cn, (f32(1) - ec) * r2


def test_shared_reciprocal_cell_matches_float64_cell():
    rng = np.random.default_rng(1)
    n = 500000
    z = rng.normal(0, 4, (4, n)).astype(f32)
    z[:, :2000] = rng.uniform(-400, 400, (4, 2000)).astype(f32)        # saturated gates (-1000 pad)
    c = rng.normal(0, 1.5, n).astype(f32)
    cn, h = _cell_shared_rcp(z[0], z[1], z[2], z[3], c, rng)
    sig = lambda v: 1 / (1 + np.exp(-np.clip(v.astype(np.float64), -700, 700)))
    cn_ref = sig(z[1]) * c.astype(np.float64) + sig(z[0]) * np.tanh(z[2].astype(np.float64))
    h_ref = sig(z[3]) * np.tanh(cn_ref)
    assert np.isfinite(cn).all() and np.isfinite(h).all()
    assert np.abs(cn - cn_ref).max() < 3e-6 and np.abs(h - h_ref).max() < 1e-6
    assert np.sqrt(np.mean((h - h_ref) ** 2)) < 1.5e-7


def test_polynomial_exp2_is_as_accurate_as_mufu():
    y = np.concatenate([np.linspace(-125, 30, 400001), np.random.default_rng(2).uniform(-20, 20, 200000)]).astype(f32)
    t = (y + f32(12582912.0)).astype(f32)
    n = (t - f32(12582912.0)).astype(f32)
    f = (y - n).astype(f32)
    assert f.min() >= -0.5 and f.max() <= 0.5
    q = np.full_like(y, f32(0.0013390866806730628))
    for coef in (0.009666373953223228, 0.055503569543361664, 0.2402234822511673, 0.6931471824645996, 1.0):
        q = (q.astype(np.float64) * f + coef).astype(f32)
    r = (q.view(np.int32) + (t.view(np.int32) << 23)).view(f32)
    assert np.abs(r / np.exp2(y.astype(np.float64)) - 1).max() < 3e-7


def test_uncertainty_triangle_contains_the_box():
    """k_scaler_head_tc decodes at (-3.1, -1.05), (3.1, -1.05), (0, 2.15) in units of the
    (scale, shift) half-widths; every point of the box [-1, 1]^2 must be a convex combination."""
    v = np.array([[-3.1, -1.05], [3.1, -1.05], [0.0, 2.15]])
    T = np.vstack([v.T, np.ones(3)])
    g = np.linspace(-1, 1, 41)
    pts = np.array([[x, y, 1.0] for x in g for y in g]).T
    lam = np.linalg.solve(T, pts)
    assert lam.min() > 0.01
