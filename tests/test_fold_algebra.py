"""The algebra behind csrc/chain.cu's weight folding (chain_pack), restated in torch and checked against the oracle
on CPU.  Three folds:

* every Linear between two LayerNorms collapses into one matrix, and the block Linear of the NEXT stage rides in the
  same contraction ([h | u] rows);
* LayerNorm_b of a stage (v2:549) is applied AFTER the next contraction: with n = r Gamma (h2 - mu) + beta,
  W n = r (W Gamma h2 - mu W Gamma 1) + W beta, so the operand is the raw h2 and the weights are [W1 | W2];
* the eps contraction of step t and the first stage of step t-1 are one phase over the operand
  [x~ | -c_b LN_f(h) | -c_b x_t] (x_{t-1} = x~ - c_b eps), whose extra rows [0 | W_f'] finish eps for the fp32 state update.

The emulation below follows the kernel's phase order (merged phase first, one tail phase after the last step) and must
reproduce ConditionalUNet.forward (v2:535-561) and ConditionalDenoiseDiffusion.sample (v2:594-598).  A second test
rounds weights and operands to bf16 to show that the 2e-2 tolerance north_star states for bf16 is met with margin."""
import pytest
import torch

from oracle import restate as R, weights

HID = [256, 512, 1024, 512, 256]
L = 256


def fold(sd, n_t=1000, dtype=torch.float64):
    """Natural-order folded phases, as chain_pack builds them."""
    g = {k: v.to(dtype) for k, v in sd.items()}
    nst = len(HID) - 1
    te = R.time_embedding(sd, torch.arange(n_t)).to(dtype)            # (n_t, 256)
    ce = R.class_embedding(sd, torch.arange(102)).to(dtype)            # (102, 256)
    def tp(i, e):
        return e @ g[f"time_projections.{i}.weight"].T + g[f"time_projections.{i}.bias"]
    T = [tp(i, te) for i in range(nst)] + [te @ g["final_time_proj.weight"].T + g["final_time_proj.bias"]]
    C = [tp(i, ce) for i in range(nst)] + [ce @ g["final_class_proj.weight"].T + g["final_class_proj.bias"]]
    Wb = [g[f"layers.{i}.0.0.weight"] for i in range(nst)]
    bb = [g[f"layers.{i}.0.0.bias"] for i in range(nst)]
    s = torch.sigmoid(g["residual_weight"])
    Wf, bf = g["final.weight"], g["final.bias"]
    Wfp, bfin = torch.cat([Wf, s * Wf], 1), (1 + s) * bf               # eps = Wfp [LN_f(h) ; x] + bfin  (v2:560-561)
    Wlp, blp = g["latent_proj.weight"], g["latent_proj.bias"]
    G0 = torch.cat([Wlp, Wb[0] @ Wlp])                                 # [h_0 | u_0] rows
    phases = [dict(kind="merged",
                   G=torch.cat([torch.cat([G0, G0 @ Wfp], 1), torch.cat([torch.zeros(L, L, dtype=dtype), Wfp], 1)]),
                   b=torch.cat([blp, Wb[0] @ blp + bb[0]]), g0b=G0 @ bfin, bfin=bfin,
                   T=torch.cat([T[0], T[0] @ Wb[0].T], 1), C=torch.cat([C[0], C[0] @ Wb[0].T], 1))]
    for i in range(nst):
        d = HID[i]
        Wv = g[f"attention_layers.{i}.in_proj_weight"][2 * d:3 * d]
        bv = g[f"attention_layers.{i}.in_proj_bias"][2 * d:3 * d]
        Wo, bo = g[f"attention_layers.{i}.out_proj.weight"], g[f"attention_layers.{i}.out_proj.bias"]
        A, a = Wo @ Wv, Wo @ bv + bo                                   # the L = 1 attention (v2:550-551)
        gam, beta = g[f"layers.{i}.1.weight"], g[f"layers.{i}.1.bias"]  # LayerNorm_b
        Wd, bd = g[f"layers.{i}.2.weight"], g[f"layers.{i}.2.bias"]
        W1, W2, b = Wd, (Wd @ A) * gam[None, :], Wd @ (A @ beta + a) + bd
        Tn, Cn = T[i + 1], C[i + 1]
        if i < nst - 1:
            j = i + 1
            W1, W2, b = torch.cat([W1, Wb[j] @ W1]), torch.cat([W2, Wb[j] @ W2]), torch.cat([b, Wb[j] @ b + bb[j]])
            Tn, Cn = torch.cat([Tn, Tn @ Wb[j].T], 1), torch.cat([Cn, Cn @ Wb[j].T], 1)
        phases.append(dict(kind="dual", G=torch.cat([W1, W2], 1), q=W2.sum(1), b=b, T=Tn, C=Cn))
    return [{k: (v.float() if isinstance(v, torch.Tensor) else v) for k, v in p.items()} for p in phases]


def run_chain(sd, phases, x, ts, c, noise=None, bf16=False):
    """The kernel's phase sequence for the timesteps `ts` (one entry: forward(); several: the reverse chain).
    Returns (eps of the last timestep, x after the last update)."""
    rnd = (lambda v: v.to(torch.bfloat16).float()) if bf16 else (lambda v: v)
    nst = len(HID) - 1
    ln = lambda v, n: torch.nn.functional.layer_norm(v, (v.shape[-1],), sd[n + ".weight"], sd[n + ".bias"], 1e-5)
    sched = R.schedule(1000)
    beta, alpha, alpha_bar = sched
    sample = noise is not None
    M = phases[0]
    nsr = 2 * HID[0]
    op = torch.cat([x, torch.zeros_like(x), torch.zeros_like(x)], 1)    # [x | 0 | 0]
    cb_prev, eps = 0.0, None
    for it in range(len(ts) + 1):
        tail = it == len(ts)
        acc = rnd(op) @ rnd(M["G"]).T                                   # merged phase: stage rows and eps rows
        if it > 0:                                                      # eps rows: -c_b (eps - b_fin) of the previous forward
            tp_ = ts[it - 1]
            eps = M["bfin"] - acc[:, nsr:] / cb_prev
            if sample:
                x = R.ddpm_update(sched, x, eps, tp_, noise[it - 1])
        if tail:
            break
        t = ts[it]
        y = acc[:, :nsr] + M["b"] + M["T"][t] - cb_prev * M["g0b"]
        if c is not None:
            y = y + M["C"][c]
        if sample:                                                      # pieces of the next merged operand
            c2, sa = float((1 - alpha[t]) / torch.sqrt(1 - alpha_bar[t])), float(torch.sqrt(alpha[t]))
            sig = float(torch.sqrt(beta[t])) if t > 0 else 0.0
            cb = c2 / sa
            xt = x / sa + sig * noise[it]
        else:
            cb, xt = 1.0, x
        for i in range(nst):
            d = HID[i]
            h, u = y[:, :d], y[:, d:]
            h2 = R.swish(ln(u, f"layers.{i}.0.1")) + h                  # v2:546-548
            mu, r = h2.mean(1, keepdim=True), 1.0 / torch.sqrt(h2.var(1, unbiased=False, keepdim=True) + 1e-5)
            P = phases[i + 1]
            # operand: raw h2; weights [W1 | W2], two accumulators
            W1h, W2h = rnd(h2) @ rnd(P["G"][:, :d]).T, rnd(h2) @ rnd(P["G"][:, d:]).T
            y = W1h + r * (W2h - mu * P["q"]) + P["b"] + P["T"][t]
            if c is not None:
                y = y + P["C"][c]
        hf = ln(y, "final_norm")
        op = torch.cat([xt, -cb * hf, -cb * x], 1)
        cb_prev = cb
    return eps, x


@pytest.mark.parametrize("kind", ["init", "perturbed"])
@pytest.mark.parametrize("with_class", [True, False])
def test_folded_phases_equal_the_reference_forward(kind, with_class):
    torch.manual_seed(0)
    sd = weights.make_unet_state(42, kind)
    phases = fold(sd)
    assert [tuple(p["G"].shape) for p in phases] == [(768, 768), (1024, 512), (2048, 1024), (1024, 2048), (256, 1024)]
    B = 8
    x = torch.randn(B, 256) * 3.0
    c = torch.randint(0, 102, (B,)) if with_class else None
    for tval in (0, 1, 500, 999):
        ref = R.unet_forward(sd, x, torch.tensor([tval]), c)
        got, _ = run_chain(sd, phases, x, [tval], c)
        assert R.max_rel(got, ref) < 5e-5, (kind, tval, R.max_rel(got, ref))


@pytest.mark.parametrize("kind", ["init", "perturbed"])
def test_merged_phase_chain_equals_the_reference_sampler(kind):
    """Four reverse steps (incl. t = 0, which adds no noise) through the merged-phase formulation vs sample() (v2:594-598)."""
    torch.manual_seed(2)
    sd = weights.make_unet_state(42, kind)
    phases = fold(sd)
    B = 6
    x = torch.randn(B, 256)
    c = torch.randint(0, 102, (B,))
    noise = [torch.randn(B, 256) for _ in range(4)]
    ts = [3, 2, 1, 0]
    want, _ = R.sample(sd, R.schedule(1000), x, c, noise_fn=lambda t: noise[3 - t], t_start=3)
    _, got = run_chain(sd, phases, x, ts, c, noise=noise)
    assert R.rel_l2(got, want) < 1e-4, R.rel_l2(got, want)


@pytest.mark.parametrize("kind", ["init", "perturbed"])
def test_bf16_operand_rounding_stays_inside_the_stated_tolerance(kind):
    torch.manual_seed(1)
    sd = weights.make_unet_state(42, kind)
    phases = fold(sd)
    x = torch.randn(16, 256) * 50.0          # the chain's x grows large under random weights
    c = torch.randint(0, 102, (16,))
    ref = R.unet_forward(sd, x, torch.tensor([321]), c)
    got, _ = run_chain(sd, phases, x, [321], c, bf16=True)
    assert R.max_rel(got, ref) < 1e-2        # north_star: 2e-2 for bf16
