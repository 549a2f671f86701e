"""The algebra behind csrc/chain.cu's weight folding (chain_pack), restated in torch and checked against the
oracle on CPU: every Linear between two LayerNorms collapses into one matrix, the Linear_b of the NEXT stage rides
in the same GEMM, and the result must still be ConditionalUNet.forward (v2:535-561).

Also emulates the bf16 path's roundings (bf16 weights and operands, fp32 accumulation and statistics) to show the
2e-2 tolerance north_star states for bf16 is met with margin by this formulation."""
import pytest
import torch

from oracle import restate as R, weights

HID = [256, 512, 1024, 512, 256]


def fold(sd, n_t=1000, dtype=torch.float64):
    """Natural-order folded phases, as chain_pack builds them: list of dict(G, b, T, C) + LN params."""
    g = {k: v.to(dtype) for k, v in sd.items()}
    nst = len(HID) - 1
    t_all = torch.arange(n_t)
    te = R.time_embedding(sd, t_all).to(dtype)                       # (n_t, 256)
    ce = R.class_embedding(sd, torch.arange(102)).to(dtype)           # (102, 256)
    def tp(i, e):
        w, b = (g[f"time_projections.{i}.weight"], g[f"time_projections.{i}.bias"]) if i < nst else (None, None)
        return e @ w.T + b
    T = [tp(i, te) for i in range(nst)] + [te @ g["final_time_proj.weight"].T + g["final_time_proj.bias"]]
    C = [tp(i, ce) for i in range(nst)] + [ce @ g["final_class_proj.weight"].T + g["final_class_proj.bias"]]
    Wb = [g[f"layers.{i}.0.0.weight"] for i in range(nst)]
    bb = [g[f"layers.{i}.0.0.bias"] for i in range(nst)]
    phases = []
    Wlp, blp = g["latent_proj.weight"], g["latent_proj.bias"]
    phases.append(dict(G=torch.cat([Wlp, Wb[0] @ Wlp]), b=torch.cat([blp, Wb[0] @ blp + bb[0]]),
                       T=torch.cat([T[0], T[0] @ Wb[0].T], 1), C=torch.cat([C[0], C[0] @ Wb[0].T], 1)))
    for i in range(nst):
        d = HID[i]
        Wv = g[f"attention_layers.{i}.in_proj_weight"][2 * d:3 * d]
        bv = g[f"attention_layers.{i}.in_proj_bias"][2 * d:3 * d]
        Wo, bo = g[f"attention_layers.{i}.out_proj.weight"], g[f"attention_layers.{i}.out_proj.bias"]
        A, a = Wo @ Wv, Wo @ bv + bo
        Wd, bd = g[f"layers.{i}.2.weight"], g[f"layers.{i}.2.bias"]
        D, db = torch.cat([Wd, Wd @ A], 1), Wd @ a + bd
        if i < nst - 1:
            j = i + 1
            phases.append(dict(G=torch.cat([D, Wb[j] @ D]), b=torch.cat([db, Wb[j] @ db + bb[j]]),
                               T=torch.cat([T[j], T[j] @ Wb[j].T], 1), C=torch.cat([C[j], C[j] @ Wb[j].T], 1)))
        else:
            phases.append(dict(G=D, b=db, T=T[nst], C=C[nst]))
    s = torch.sigmoid(g["residual_weight"])
    Wf, bf = g["final.weight"], g["final.bias"]
    phases.append(dict(G=torch.cat([Wf, s * Wf], 1), b=(1 + s) * bf, T=None, C=None))
    return [{k: (v.float() if v is not None else None) for k, v in p.items()} for p in phases]


def run_folded(sd, phases, x, t, c, bf16=False):
    """One denoiser evaluation through the folded phases; bf16=True rounds weights and GEMM operands to bf16."""
    rnd = (lambda v: v.to(torch.bfloat16).float()) if bf16 else (lambda v: v)
    nst = len(HID) - 1
    def gemm(p, a):
        y = rnd(a) @ rnd(p["G"]).T + p["b"]
        if p["T"] is not None:
            y = y + p["T"][t]
            if c is not None:
                y = y + p["C"][c]
        return y
    ln = lambda v, n: torch.nn.functional.layer_norm(v, (v.shape[-1],), sd[n + ".weight"], sd[n + ".bias"], 1e-5)
    y = gemm(phases[0], x)
    for i in range(nst):
        d = HID[i]
        h, u = y[:, :d], y[:, d:]
        h2 = R.swish(ln(u, f"layers.{i}.0.1")) + h
        n = ln(h2, f"layers.{i}.1")
        y = gemm(phases[i + 1], torch.cat([h2, n], 1))
    hf = ln(y, "final_norm")
    return gemm(phases[nst + 1], torch.cat([hf, x], 1))


@pytest.mark.parametrize("kind", ["init", "perturbed"])
@pytest.mark.parametrize("with_class", [True, False])
def test_folded_phases_equal_the_reference_forward(kind, with_class):
    torch.manual_seed(0)
    sd = weights.make_unet_state(42, kind)
    phases = fold(sd)
    assert [tuple(p["G"].shape) for p in phases] == [(512, 256), (1024, 512), (2048, 1024), (1024, 2048), (256, 1024), (256, 512)]
    B = 8
    x = torch.randn(B, 256) * 3.0
    c = torch.randint(0, 102, (B,)) if with_class else None
    for tval in (0, 1, 500, 999):
        t = torch.tensor([tval])
        ref = R.unet_forward(sd, x, t, c)
        got = run_folded(sd, phases, x, t, c)
        assert R.max_rel(got, ref) < 2e-5, (kind, tval)


@pytest.mark.parametrize("kind", ["init", "perturbed"])
def test_bf16_operand_rounding_stays_inside_the_stated_tolerance(kind):
    torch.manual_seed(1)
    sd = weights.make_unet_state(42, kind)
    phases = fold(sd)
    x = torch.randn(16, 256) * 50.0          # the chain's x grows large under random weights
    c = torch.randint(0, 102, (16,))
    t = torch.tensor([321])
    ref = R.unet_forward(sd, x, t, c)
    got = run_folded(sd, phases, x, t, c, bf16=True)
    assert R.max_rel(got, ref) < 1e-2        # north_star: 2e-2 for bf16
