"""GPU tests of the input guards and host-side contracts around the kernels: invalid indices and padding lengths are
reported (the reference raises IndexError at the same places), cached weight planes can be invalidated, evaluation
metrics follow utils.test / utils.test_phonon on degenerate targets, the fused AdamW keeps per-parameter step counts."""
import pytest
import torch

from dostransformer_b200 import _lib as L
from dostransformer_b200 import ops
from dostransformer_b200.synthetic import make_edos_batch
from oracle import dost_oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda"


def test_out_of_range_index_is_skipped_and_reported():
    """DOSTransformer.py:139-140 / :187 index with edge_index; an index past the node table raises there.  Here the
    kernel skips it (no out-of-bounds write) and the next poll raises."""
    L.lib()
    torch.cuda.synchronize()
    L.poll_device_errors()
    key = torch.tensor([0, 3, 2, 9, 1, -1, 3], dtype=torch.int32, device=DEV)      # 9 and -1 are outside [0, 4)
    csr, _ = ops.csr_build(key, 4)
    torch.cuda.synchronize()
    with pytest.raises(RuntimeError, match="outside its table"):
        L.poll_device_errors()
    L.poll_device_errors()                      # the poll cleared the flag
    # the valid keys still form a correct CSR (segments 0..3 hold elements 0 | 4 | 2 | 1, 6)
    assert csr.rowptr.cpu().tolist() == [0, 1, 2, 3, 5]
    assert csr.perm.cpu().tolist()[:5] == [0, 4, 2, 1, 6]


def test_stale_padding_length_is_rejected():
    """A data-parallel padding length smaller than the batch's own largest crystal (a stale model.max_num_nodes) must not
    reach the kernels."""
    g = make_edos_batch(5, seed=3, mean_atoms=10.0).to(DEV)
    with pytest.raises(ValueError, match="stale data-parallel padding length"):
        ops.build_graph(g.edge_index, g.batch, g.system, nmax_override=g.max_num_nodes - 1, nmax_hint=g.max_num_nodes)
    # without a host hint the device value becomes max(override, measured): never below the batch's own maximum
    # (and the padding length is read back once, so that the batch still takes the host-sized tensor-core attention)
    gr = ops.build_graph(g.edge_index, g.batch, g.system, nmax_override=2)
    assert int(gr.nmax.item()) == g.max_num_nodes == gr.nmax_host
    gr = ops.build_graph(g.edge_index, g.batch, g.system, nmax_override=g.max_num_nodes + 7, nmax_hint=g.max_num_nodes)
    assert int(gr.nmax.item()) == g.max_num_nodes + 7 == gr.nmax_host


def test_weight_plane_cache_invalidation():
    w = torch.nn.Parameter(torch.randn(64, 64, device=DEV))
    with ops.precision("bf16x3"):
        a = ops.weight_planes(w)
        assert ops.weight_planes(w) is a                       # cache hit
        with torch.no_grad():
            w.mul_(2.0)                                        # version bump: seen
        b = ops.weight_planes(w)
        assert b is not a and torch.equal(b.hi.float(), (w.detach()).bfloat16().float())
        w.data.mul_(0.5)                                       # bypasses the version counter: NOT seen ...
        assert ops.weight_planes(w) is b
        ops.invalidate_weight_planes([w])                      # ... until invalidated
        c = ops.weight_planes(w)
        assert c is not b and torch.equal(c.hi.float(), w.detach().bfloat16().float())
        ops.invalidate_weight_planes()                         # global generation bump
        assert ops.weight_planes(w) is not c


def test_eval_metrics_degenerate_targets_and_phonon_mode():
    B, T = 4, 51
    gen = torch.Generator().manual_seed(0)
    pred = torch.randn(B, T, generator=gen).to(DEV)
    y = torch.randn(B, T, generator=gen)
    y[0] = -0.3                    # clamps to an all-zero (constant) target in eDOS mode
    y[1] = 0.7                     # constant target
    pred[1] = 0.7                  # ... predicted exactly
    y = y.to(DEV)
    per, mean = ops.eval_metrics(pred, y, clamp_pred=True)
    per = per.cpu()
    assert torch.isfinite(per).all()
    assert per[0, 3].item() == 0.0 and per[1, 3].item() == 1.0          # sklearn r2_score on a constant target
    yc, pc = y.clamp_min(0).cpu().double(), pred.clamp_min(0).cpu().double()
    for b in (2, 3):
        want = O.eval_metrics(pc[b:b + 1], yc[b:b + 1], clamp_pred=True)
        assert abs(per[b, 0].item() - want["mse"].item()) < 1e-5 * max(1.0, want["mse"].item())
        assert abs(per[b, 3].item() - want["r2"].item()) < 1e-4 * max(1.0, abs(want["r2"].item()))
    # phonon mode (utils.py:127-131): neither the target nor the prediction is clamped
    per_p, _ = ops.eval_metrics(pred, y, clamp_pred=False)
    per_p = per_p.cpu()
    for b in (2, 3):
        want = O.eval_metrics(pred[b:b + 1].cpu().double(), y[b:b + 1].cpu().double(), clamp_pred=False)
        assert abs(per_p[b, 0].item() - want["mse"].item()) < 1e-5 * max(1.0, want["mse"].item())
        assert abs(per_p[b, 2].item() - want["mae"].item()) < 1e-5 * max(1.0, want["mae"].item())
        assert abs(per_p[b, 3].item() - want["r2"].item()) < 1e-4 * max(1.0, abs(want["r2"].item()))


def test_fused_adamw_per_parameter_step_and_param_groups():
    """torch.optim.AdamW keeps one step count per parameter: a parameter without a gradient in the first steps gets its
    own bias correction later.  Param-group dicts with their own lr are honoured."""
    from dostransformer_b200.optim import AdamW
    torch.manual_seed(0)
    ps = [torch.nn.Parameter(torch.randn(33, 17, device=DEV)), torch.nn.Parameter(torch.randn(129, device=DEV)),
          torch.nn.Parameter(torch.randn(8, 8, device=DEV))]
    qs = [torch.nn.Parameter(p.detach().clone()) for p in ps]
    mine = AdamW([{"params": ps[:2]}, {"params": ps[2:], "lr": 3e-3}], lr=1e-3, weight_decay=1e-2)
    ref = torch.optim.AdamW([{"params": qs[:2]}, {"params": qs[2:], "lr": 3e-3}], lr=1e-3, weight_decay=1e-2)
    gen = torch.Generator().manual_seed(1)
    for it in range(5):
        for i, (p, q) in enumerate(zip(ps, qs)):
            if i == 1 and it < 2:              # no gradient for the second parameter in the first two steps
                p.grad = q.grad = None
                continue
            g = torch.randn(p.shape, generator=gen).to(DEV)
            p.grad, q.grad = g.clone(), g.clone()
        mine.step()
        ref.step()
    for p, q in zip(ps, qs):
        assert (p - q).abs().max().item() <= 1e-6 * max(1.0, q.abs().max().item())
    assert mine.state[ps[1]]["step"] == 3 and mine.state[ps[0]]["step"] == 5
    sd = mine.state_dict()
    assert [g["lr"] for g in sd["param_groups"]] == [1e-3, 3e-3] and sd["param_groups"][1]["params"] == [2]
    with pytest.raises(ValueError):
        AdamW([{"params": ps[:1], "momentum": 0.9}])


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_model_on_a_non_current_device():
    """The model makes its own device current for the call; raw ops refuse tensors of another device."""
    from dostransformer_b200.embedder_eDOS.DOSTransformer import DOSTransformer
    dev1 = torch.device("cuda", 1)
    torch.manual_seed(0)
    m0 = DOSTransformer(1, 1, 200, 41, 2, 128, torch.device("cuda", 0), 0.0).to("cuda:0")
    torch.manual_seed(0)
    m1 = DOSTransformer(1, 1, 200, 41, 2, 128, dev1, 0.0).to(dev1)
    g = make_edos_batch(4, seed=5, mean_atoms=8.0)
    assert torch.cuda.current_device() == 0
    a = m0(g.clone().to("cuda:0"))[0]
    b = m1(g.clone().to(dev1))[0]                   # current device is still 0 here
    assert b.device == dev1 and torch.equal(a.cpu(), b.cpu())
    with pytest.raises(RuntimeError, match="current device"):
        ops.split_planes(torch.randn(8, 8, device=dev1))


def test_batch_without_padding_hint_takes_the_tensor_core_attention(monkeypatch):
    """A stock PyG Batch (main_eDOS.py:54) has no max_num_nodes: build_graph reads the padding length back once and the
    model runs the same kernels - bitwise the same outputs - as with the collate's hint; DOST_NO_NMAX_SYNC=1 keeps the
    step free of device->host reads and falls back to the FMA-pipe attention (same results to the bf16x3 floor)."""
    from dostransformer_b200.embedder_eDOS.DOSTransformer import DOSTransformer
    torch.manual_seed(0)
    m = DOSTransformer(3, 2, 200, 41, 2, 128, torch.device(DEV), 0.0, precision="bf16x3").to(DEV).eval()
    g = make_edos_batch(6, seed=5, mean_atoms=9.0).to(DEV)
    with torch.no_grad():
        want = m(g)[0]
        hint = g.max_num_nodes
        del g.max_num_nodes
        g._keys.remove("max_num_nodes")
        assert getattr(g, "max_num_nodes", None) is None
        got = m(g)[0]
        assert torch.equal(got, want)
        gr = ops.build_graph(g.edge_index, g.batch, g.system)
        assert gr.nmax_host == hint
        monkeypatch.setenv("DOST_NO_NMAX_SYNC", "1")
        L.reload_switches()
        try:
            assert ops.build_graph(g.edge_index, g.batch, g.system).nmax_host is None
            fma = m(g)[0]
        finally:
            monkeypatch.delenv("DOST_NO_NMAX_SYNC")
            L.reload_switches()
        assert (fma - want).abs().max() / want.abs().max() < 1e-4
