"""GPU parity tests of the individual kernels (through the C ABI via dostransformer_b200.ops) against plain
torch restatements.  Integer work is bit-exact; floating point tolerances are written at each assert."""
import math

import pytest
import torch

from conftest import relerr
from dostransformer_b200 import _lib as L
from dostransformer_b200 import ops
from dostransformer_b200.ops import RowMap
from oracle import dost_oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda"
TOL = {torch.float32: 2e-5, torch.float64: 1e-12}


def _rand(*shape, dtype=torch.float32, seed=0, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(*shape, generator=g, dtype=torch.float64) * scale).to(dtype).to(DEV)


# --------------------------------------------------------------------------------------------- integer structure
@pytest.mark.parametrize("n,size,hub", [(1000, 37, False), (5000, 300, True), (12, 7, False), (1, 5, False),
                                        (100000, 9000, True)])
def test_csr_build_bit_exact(n, size, hub):
    g = torch.Generator().manual_seed(n)
    key = torch.randint(0, size, (n,), generator=g)
    if hub:
        key[torch.rand(n, generator=g) < 0.3] = size // 2       # a hub destination (index-0 padding in mat2graph.py)
        key[key == 3] = 4                                        # and an empty segment
    rowptr_ref, perm_ref = O.csr_by_key(key, size)
    csr, mx = ops.csr_build(ops.to_i32(key.to(DEV)), size, want_max=True)
    assert torch.equal(csr.rowptr.cpu().long(), rowptr_ref)
    assert torch.equal(csr.perm.cpu().long(), perm_ref)
    assert int(mx.item()) == int(torch.bincount(key, minlength=size).max())


def test_build_graph_matches_oracle():
    from dostransformer_b200.synthetic import make_edos_batch
    g = make_edos_batch(9, seed=11)
    gr = ops.build_graph(g.edge_index.to(DEV), g.batch.to(DEV), g.system.to(DEV))
    ptr, nmax = O.crystal_ptr(g.batch)
    assert torch.equal(gr.ptr.cpu().long(), ptr) and int(gr.nmax.item()) == nmax
    rp, pm = O.csr_by_key(g.edge_index[1], g.batch.numel())
    assert torch.equal(gr.by_dst.rowptr.cpu().long(), rp) and torch.equal(gr.by_dst.perm.cpu().long(), pm)
    rp, pm = O.csr_by_key(g.edge_index[0], g.batch.numel())
    assert torch.equal(gr.by_src.rowptr.cpu().long(), rp) and torch.equal(gr.by_src.perm.cpu().long(), pm)
    rp, pm = O.csr_by_key(g.system, 7)
    assert torch.equal(gr.by_system.rowptr.cpu().long(), rp) and torch.equal(gr.by_system.perm.cpu().long(), pm)


# --------------------------------------------------------------------------------------------- GEMM
@pytest.mark.parametrize("dtype", [torch.float32, torch.float64])
@pytest.mark.parametrize("M,N,K", [(300, 70, 41), (129, 257, 200), (5, 1, 64), (1024, 512, 768), (64, 64, 2)])
def test_gemm_plain_and_epilogues(dtype, M, N, K):
    a, w, b = _rand(M, K, dtype=dtype, seed=1), _rand(N, K, dtype=dtype, seed=2), _rand(N, dtype=dtype, seed=3)
    res = _rand(M, N, dtype=dtype, seed=4)
    out = torch.empty(M, N, dtype=dtype, device=DEV)
    pre = torch.empty(M, N, dtype=dtype, device=DEV)
    slope = torch.tensor([0.25], dtype=dtype, device=DEV)
    ops.gemm_raw(M=M, N=N, K=K, a=[(a, None)], a_mode=L.KC, b=w, b_mode=L.KC, out=out, bias=b, act=L.ACT_PRELU,
                 prelu_slope=slope, out_pre=pre, residual=res)
    v = a.double() @ w.double().T + b.double()
    ref = torch.nn.functional.prelu(v, slope.double()) + res.double()
    assert relerr(pre, v) < TOL[dtype] * 5
    assert relerr(out, ref) < TOL[dtype] * 5
    ops.gemm_raw(M=M, N=N, K=K, a=[(a, None)], a_mode=L.KC, b=w, b_mode=L.KC, out=out, act=L.ACT_RELU)
    assert relerr(out, torch.relu(a.double() @ w.double().T)) < TOL[dtype] * 5
    ops.gemm_raw(M=M, N=N, K=K, a=[(a, None)], a_mode=L.KC, b=w, b_mode=L.KC, out=out, act=L.ACT_LEAKY, act_slope=0.01,
                 accumulate=True)
    ref2 = torch.relu(a.double() @ w.double().T) + torch.nn.functional.leaky_relu(a.double() @ w.double().T, 0.01)
    assert relerr(out, ref2) < TOL[dtype] * 5


@pytest.mark.parametrize("dtype", [torch.float32, torch.float64])
def test_gemm_modes_and_splitk(dtype):
    M, N, K = 777, 96, 130
    dy, x, w = _rand(M, N, dtype=dtype, seed=1), _rand(M, K, dtype=dtype, seed=2), _rand(N, K, dtype=dtype, seed=3)
    # dX = dY @ W  (B read N-contiguous)
    dx = torch.empty(M, K, dtype=dtype, device=DEV)
    ops.gemm_raw(M=M, N=K, K=N, a=[(dy, None)], a_mode=L.KC, b=w, b_mode=L.MC, out=dx)
    assert relerr(dx, dy.double() @ w.double()) < TOL[dtype] * 5
    # dW = dY^T @ X with deterministic split-K, written into a column slice of a wider matrix
    dw_full = torch.zeros(N, K + 32, dtype=dtype, device=DEV)
    for split in (1, 5):
        ops.gemm_raw(M=N, N=K, K=M, a=[(dy, None)], a_mode=L.MC, b=x, b_mode=L.MC, out=dw_full[:, 32:], ldc=K + 32,
                     split_k=split)
        assert relerr(dw_full[:, 32:], dy.double().T @ x.double()) < TOL[dtype] * 10
        assert dw_full[:, :32].abs().max() == 0
    # gathered reduction rows: dW = dY^T @ X[idx]
    idx = torch.randint(0, 50, (M,), dtype=torch.int32, device=DEV)
    xs = _rand(50, K, dtype=dtype, seed=5)
    dw = torch.empty(N, K, dtype=dtype, device=DEV)
    ops.gemm_raw(M=N, N=K, K=M, a=[(dy, None)], a_mode=L.MC, b=xs, b_mode=L.MC, b_map=RowMap(idx=idx), out=dw, split_k=3)
    assert relerr(dw, dy.double().T @ xs.double()[idx.long()]) < TOL[dtype] * 10


@pytest.mark.parametrize("dtype", [torch.float32, torch.float64])
def test_gemm_gather_concat_and_batched(dtype):
    E, Nn, H = 500, 60, 32
    x, e = _rand(Nn, H, dtype=dtype, seed=1), _rand(E, H, dtype=dtype, seed=2)
    w = _rand(2 * H, 3 * H, dtype=dtype, seed=3)
    row = torch.randint(0, Nn, (E,), dtype=torch.int32, device=DEV)
    col = torch.randint(0, Nn, (E,), dtype=torch.int32, device=DEV)
    out = torch.empty(E, 2 * H, dtype=dtype, device=DEV)
    ops.gemm_raw(M=E, N=2 * H, K=3 * H, a=[(x, RowMap(idx=row)), (x, RowMap(idx=col)), (e, None)], a_mode=L.KC, b=w,
                 b_mode=L.KC, out=out)
    cat = torch.cat([x[row.long()], x[col.long()], e], 1).double()
    assert relerr(out, cat @ w.double().T) < TOL[dtype] * 5
    # per-crystal broadcast rows and an index + broadcast map (fc_prompt shape)
    B, T = 7, 11
    en, gv, tab = _rand(B * T, H, dtype=dtype, seed=4), _rand(B, H, dtype=dtype, seed=5), _rand(7, 16, dtype=dtype, seed=6)
    sysid = torch.randint(0, 7, (B,), dtype=torch.int32, device=DEV)
    w2 = _rand(H, 2 * H + 16, dtype=dtype, seed=7)
    out2 = torch.empty(B * T, H, dtype=dtype, device=DEV)
    ops.gemm_raw(M=B * T, N=H, K=2 * H + 16, a=[(en, None), (gv, RowMap(div=T)), (tab, RowMap(idx=sysid, div=T))],
                 a_mode=L.KC, b=w2, b_mode=L.KC, out=out2)
    cat2 = torch.cat([en, gv.repeat_interleave(T, 0), tab[sysid.long()].repeat_interleave(T, 0)], 1).double()
    assert relerr(out2, cat2 @ w2.double().T) < TOL[dtype] * 5
    # batched NT with padded ldc
    S, T2 = 5, 13
    q, k = _rand(S, T2, H, dtype=dtype, seed=8), _rand(S, T2, H, dtype=dtype, seed=9)
    sc = torch.zeros(S, T2, 16, dtype=dtype, device=DEV)
    ops.gemm_raw(M=T2, N=T2, K=H, a=[(q.view(-1, H), None)], a_mode=L.KC, b=k.view(-1, H), b_mode=L.KC, out=sc, batch=S,
                 a_bstride=T2 * H, b_bstride=T2 * H, c_bstride=T2 * 16, ldc=16)
    assert relerr(sc[:, :, :T2], torch.bmm(q.double(), k.double().transpose(1, 2))) < TOL[dtype] * 5
    assert sc[:, :, T2:].abs().max() == 0


# --------------------------------------------------------------------------------------------- autograd ops
def _grads(fn, inputs):
    for t in inputs:
        t.grad = None
    out = fn()
    outs = out if isinstance(out, (tuple, list)) else [out]
    g = torch.Generator().manual_seed(99)
    total = 0
    for o in outs:
        wgt = torch.randn(o.shape, generator=g, dtype=torch.float64).to(o.dtype).to(o.device)
        total = total + (o * wgt).sum()
    total.backward()
    return [o.detach() for o in outs], [t.grad.detach().clone() if t.grad is not None else None for t in inputs]


def _leaf(t):
    return t.detach().clone().requires_grad_(True)


@pytest.mark.parametrize("dtype", [torch.float32, torch.float64])
def test_linear_autograd_edge_mlp_shape(dtype):
    from dostransformer_b200.synthetic import make_edos_batch
    g = make_edos_batch(5, seed=21, mean_atoms=6.0)
    gr = ops.build_graph(g.edge_index.to(DEV), g.batch.to(DEV), g.system.to(DEV))
    H = 32
    x, e = _leaf(_rand(gr.N, H, dtype=dtype, seed=1)), _leaf(_rand(gr.E, H, dtype=dtype, seed=2))
    w, b = _leaf(_rand(2 * H, 3 * H, dtype=dtype, seed=3, scale=0.2)), _leaf(_rand(2 * H, dtype=dtype, seed=4))
    res = _leaf(_rand(gr.E, 2 * H, dtype=dtype, seed=5))
    row, col = g.edge_index[0].to(DEV), g.edge_index[1].to(DEV)

    def mine():
        segs = [(x, RowMap(idx=gr.row, csr=gr.by_src)), (x, RowMap(idx=gr.col, csr=gr.by_dst)), (e, None)]
        return ops.linear(segs, w, b, M=gr.E, residual=res, want_pre=True)

    def ref():
        v = torch.nn.functional.linear(torch.cat([x[row], x[col], e], 1), w, b)
        return v + res, v

    o1, g1 = _grads(mine, [x, e, w, b, res])
    o2, g2 = _grads(ref, [x, e, w, b, res])
    for a_, b_ in zip(o1 + g1, o2 + g2):
        assert relerr(a_, b_) < TOL[dtype] * 20


@pytest.mark.parametrize("dtype", [torch.float32, torch.float64])
@pytest.mark.parametrize("act", ["prelu", "relu", "leaky", "none"])
def test_linear_autograd_activations(dtype, act):
    M, K, N = 333, 41, 48
    x, w, b = _leaf(_rand(M, K, dtype=dtype, seed=1)), _leaf(_rand(N, K, dtype=dtype, seed=2)), _leaf(_rand(N, dtype=dtype, seed=3))
    slope = _leaf(torch.tensor([0.25], dtype=dtype, device=DEV))
    kw = {"prelu": dict(act=L.ACT_PRELU, prelu_slope=slope), "relu": dict(act=L.ACT_RELU),
          "leaky": dict(act=L.ACT_LEAKY, act_slope=0.01), "none": {}}[act]
    F = torch.nn.functional
    reff = {"prelu": lambda v: F.prelu(v, slope), "relu": F.relu, "leaky": lambda v: F.leaky_relu(v, 0.01),
            "none": lambda v: v}[act]
    ins = [x, w, b] + ([slope] if act == "prelu" else [])
    o1, g1 = _grads(lambda: ops.linear([(x, None)], w, b, **kw), ins)
    o2, g2 = _grads(lambda: reff(F.linear(x, w, b)), ins)
    for a_, b_ in zip(o1 + g1, o2 + g2):
        assert relerr(a_, b_) < TOL[dtype] * 20


@pytest.mark.parametrize("dtype", [torch.float32, torch.float64])
def test_linear_autograd_broadcast_maps(dtype):
    from dostransformer_b200.synthetic import make_edos_batch
    g = make_edos_batch(6, seed=22, mean_atoms=4.0)
    gr = ops.build_graph(g.edge_index.to(DEV), g.batch.to(DEV), g.system.to(DEV))
    B, T, H = gr.B, 9, 32
    en, gv = _leaf(_rand(B * T, H, dtype=dtype, seed=1)), _leaf(_rand(B, H, dtype=dtype, seed=2))
    tab = _leaf(_rand(7, H // 2, dtype=dtype, seed=3))
    w, b = _leaf(_rand(H, 2 * H + H // 2, dtype=dtype, seed=4, scale=0.2)), _leaf(_rand(H, dtype=dtype, seed=5))
    sysid = g.system.to(DEV)

    def mine():
        pc = RowMap(div=T, div_rowptr=gr.token_rowptr(T))
        ps = RowMap(idx=gr.system, div=T, div_rowptr=gr.token_rowptr(T), csr=gr.by_system)
        return ops.linear([(en, None), (gv, pc), (tab, ps)], w, b, M=B * T, act=L.ACT_LEAKY, act_slope=0.01)

    def ref():
        cat = torch.cat([en, gv.repeat_interleave(T, 0), tab[sysid].repeat_interleave(T, 0)], 1)
        return torch.nn.functional.leaky_relu(torch.nn.functional.linear(cat, w, b), 0.01)

    o1, g1 = _grads(mine, [en, gv, tab, w, b])
    o2, g2 = _grads(ref, [en, gv, tab, w, b])
    for a_, b_ in zip(o1 + g1, o2 + g2):
        assert relerr(a_, b_) < TOL[dtype] * 20


@pytest.mark.parametrize("dtype", [torch.float32, torch.float64])
@pytest.mark.parametrize("W,prelu", [(32, False), (256, False), (512, True), (64, True), (1024, False)])
def test_layer_norm(dtype, W, prelu):
    M = 517
    x = _leaf(_rand(M, W, dtype=dtype, seed=1, scale=2.0))
    gam, bet = _leaf(_rand(W, dtype=dtype, seed=2)), _leaf(_rand(W, dtype=dtype, seed=3))
    slope = _leaf(torch.tensor([0.25], dtype=dtype, device=DEV)) if prelu else None
    ins = [x, gam, bet] + ([slope] if prelu else [])
    F = torch.nn.functional

    def ref():
        y = F.layer_norm(x, (W,), gam, bet, 1e-5)
        return F.prelu(y, slope) if prelu else y

    o1, g1 = _grads(lambda: ops.layer_norm(x, gam, bet, slope), ins)
    o2, g2 = _grads(ref, ins)
    for a_, b_ in zip(o1 + g1, o2 + g2):
        assert relerr(a_, b_) < TOL[dtype] * 20


@pytest.mark.parametrize("dtype", [torch.float32, torch.float64])
@pytest.mark.parametrize("mean", [False, True])
@pytest.mark.parametrize("E,N", [(4000, 300), (30000, 2500)])     # block-per-segment kernel (<= 1024 segments) / warp-per-segment
def test_segment_reduce(dtype, mean, E, N):
    W = 256
    src = _leaf(_rand(E, W, dtype=dtype, seed=1))
    key = torch.randint(0, N - 5, (E,), generator=torch.Generator().manual_seed(3))
    key[:600] = 7                                                    # hub
    csr, _ = ops.csr_build(ops.to_i32(key.to(DEV)), N)
    kd = key.to(DEV)
    o1, g1 = _grads(lambda: ops.segment_reduce(src, csr, mean), [src])
    o2, g2 = _grads(lambda: (O.segment_mean if mean else O.segment_sum)(src, kd, N), [src])
    assert relerr(o1[0], o2[0]) < TOL[dtype] * 20 and relerr(g1[0], g2[0]) < TOL[dtype] * 20
    if not mean:   # same summation order as index_add_ on the CPU (ascending edge id): compare with a CPU run
        cpu = O.segment_sum(src.detach().cpu(), key, N)
        assert relerr(o1[0], cpu) < TOL[dtype]
    # odd width / strided input path
    src2 = _rand(E, 41, dtype=dtype, seed=2)
    out2 = ops.segment_reduce_raw(src2[:, 3:40], csr.rowptr, csr.perm, N)
    assert relerr(out2, O.segment_sum(src2[:, 3:40], kd, N)) < TOL[dtype] * 20
    # contiguous groups (no permutation), accumulating into an existing tensor: the adjoint of a per-crystal broadcast
    G, Tn = (N // 10), 201
    src3 = _rand(G * Tn, W, dtype=dtype, seed=4)
    rp = (torch.arange(G + 1, dtype=torch.int32, device=DEV) * Tn).contiguous()
    acc = _rand(G, W, dtype=dtype, seed=5)
    want = acc.double() + src3.double().view(G, Tn, W).sum(1)
    ops.segment_reduce_raw(src3, rp, None, G, out=acc, accumulate=True)
    assert relerr(acc, want) < TOL[dtype] * 20


def _dense_cross_ref(q, x_nodes, gam, bet, batch, H):
    """The reference's padded formulation: to_dense_batch zero padding, LN on q and on padded keys, attention."""
    dense, n, nmax = O.pad_crystals(x_nodes, batch)
    F = torch.nn.functional
    ln = lambda t: F.layer_norm(t, (H,), gam, bet, 1e-5)
    return q + O.attention(ln(q), ln(dense))


@pytest.mark.parametrize("dtype", [torch.float32, torch.float64])
@pytest.mark.parametrize("H,broadcast", [(32, False), (64, True), (256, False)])
def test_cross_attention_matches_padded_reference(dtype, H, broadcast):
    from dostransformer_b200.synthetic import make_edos_batch
    g = make_edos_batch(6, seed=31, mean_atoms=9.0, max_atoms=50)
    gr = ops.build_graph(g.edge_index.to(DEV), g.batch.to(DEV), g.system.to(DEV))
    B, T = gr.B, 37
    xn = _leaf(_rand(gr.N, H, dtype=dtype, seed=1))
    q = _leaf(_rand(T, H, dtype=dtype, seed=2)) if broadcast else _leaf(_rand(B, T, H, dtype=dtype, seed=2))
    gam, bet = _leaf(_rand(H, dtype=dtype, seed=3)), _leaf(_rand(H, dtype=dtype, seed=4))
    bd = g.batch.to(DEV)

    def mine():
        kv = ops.layer_norm(xn, gam, bet)
        ql = ops.layer_norm(q, gam, bet)
        return ops.cross_attention(ql, kv, bet, q, gr, B)

    def ref():
        qq = q[None].expand(B, T, H) if broadcast else q
        return _dense_cross_ref(qq, xn, gam, bet, bd, H)

    o1, g1 = _grads(mine, [xn, q, gam, bet])
    o2, g2 = _grads(ref, [xn, q, gam, bet])
    tol = 5e-5 if dtype == torch.float32 else 5e-6        # fp32 softmax inside both paths
    for a_, b_ in zip(o1 + g1, o2 + g2):
        assert relerr(a_, b_) < tol


@pytest.mark.parametrize("dtype", [torch.float32, torch.float64])
def test_cross_attention_two_sequences_per_crystal(dtype):
    from dostransformer_b200.synthetic import make_edos_batch
    g = make_edos_batch(4, seed=32, mean_atoms=5.0)
    gr = ops.build_graph(g.edge_index.to(DEV), g.batch.to(DEV), g.system.to(DEV))
    B, T, H = gr.B, 20, 32
    kv = _leaf(_rand(gr.N, H, dtype=dtype, seed=1))
    ph = _leaf(_rand(H, dtype=dtype, seed=2))
    q = _leaf(_rand(2 * B, T, H, dtype=dtype, seed=3))
    o1, g1 = _grads(lambda: ops.cross_attention(q, kv, ph, q, gr, 2 * B), [kv, ph, q])

    def ref():
        dense, n, nmax = O.pad_crystals(kv, g.batch.to(DEV))
        mask = torch.arange(nmax, device=DEV)[None, :] >= n[:, None]
        dense = torch.where(mask[:, :, None], ph[None, None, :], dense)
        dense2 = torch.cat([dense, dense], 0)
        return q + O.attention(q, dense2)

    o2, g2 = _grads(ref, [kv, ph, q])
    tol = 5e-5 if dtype == torch.float32 else 5e-6
    for a_, b_ in zip(o1 + g1, o2 + g2):
        assert relerr(a_, b_) < tol


@pytest.mark.parametrize("dtype", [torch.float32, torch.float64])
@pytest.mark.parametrize("S,Lq,Lk,H", [(3, 51, 51, 32), (2, 201, 201, 256), (4, 17, 40, 64)])
def test_dense_attention(dtype, S, Lq, Lk, H):
    q, k = _leaf(_rand(S, Lq, H, dtype=dtype, seed=1)), _leaf(_rand(S, Lk, H, dtype=dtype, seed=2))
    r = _leaf(_rand(S, Lq, H, dtype=dtype, seed=3))
    o1, g1 = _grads(lambda: ops.self_attention(q, k, r), [q, k, r])
    o2, g2 = _grads(lambda: r + O.attention(q, k), [q, k, r])
    tol = 5e-5 if dtype == torch.float32 else 5e-6
    for a_, b_ in zip(o1 + g1, o2 + g2):
        assert relerr(a_, b_) < tol
    if Lq == Lk:    # same tensor as query and key (first self-attention layer)
        o1, g1 = _grads(lambda: ops.self_attention(q, q, r), [q, r])
        o2, g2 = _grads(lambda: r + O.attention(q, q), [q, r])
        for a_, b_ in zip(o1 + g1, o2 + g2):
            assert relerr(a_, b_) < tol


def test_attention_dropout_properties():
    from dostransformer_b200.synthetic import make_edos_batch
    g = make_edos_batch(4, seed=33, mean_atoms=12.0)
    gr = ops.build_graph(g.edge_index.to(DEV), g.batch.to(DEV), g.system.to(DEV))
    B, T, H = gr.B, 64, 32
    kv, ph, q = _rand(gr.N, H, seed=1), _rand(H, seed=2), _rand(B, T, H, seed=3)
    zero = torch.zeros_like(q)
    base = ops.cross_attention(q, kv, ph, zero, gr, B, 0.0, 0)
    a = ops.cross_attention(q, kv, ph, zero, gr, B, 0.3, 1234)
    b = ops.cross_attention(q, kv, ph, zero, gr, B, 0.3, 1234)
    assert torch.equal(a, b)                                     # deterministic per seed
    acc = torch.zeros_like(base)
    n = 200
    for s in range(n):
        acc += ops.cross_attention(q, kv, ph, zero, gr, B, 0.3, 5000 + s)
    assert relerr(acc / n, base) < 0.15                          # inverted dropout is unbiased
    S, L_ = 3, 40
    qq, kk = _rand(S, L_, H, seed=4), _rand(S, L_, H, seed=5)
    z = torch.zeros_like(qq)
    base = ops.self_attention(qq, kk, z)
    acc = torch.zeros_like(base)
    for s in range(n):
        acc += ops.self_attention(qq, kk, z, 0.3, 7000 + s)
    assert relerr(acc / n, base) < 0.15
    # dropout backward is consistent with its forward (finite differences on a tiny case, fp64)
    q64 = _leaf(_rand(1, 3, H, dtype=torch.float64, seed=6))
    k64 = _rand(1, 5, H, dtype=torch.float64, seed=7)
    out = ops.self_attention(q64, k64, torch.zeros_like(q64), 0.4, 42)
    wgt = _rand(1, 3, H, dtype=torch.float64, seed=8)
    (out * wgt).sum().backward()
    eps = 1e-3          # the softmax inside is fp32 (reference quirk): finite differences need a coarse step
    d = torch.zeros_like(q64)
    d[0, 1, 3] = eps
    with torch.no_grad():
        f1 = (ops.self_attention(q64 + d, k64, torch.zeros_like(q64), 0.4, 42) * wgt).sum()
        f0 = (ops.self_attention(q64 - d, k64, torch.zeros_like(q64), 0.4, 42) * wgt).sum()
    assert abs(((f1 - f0) / (2 * eps)).item() - q64.grad[0, 1, 3].item()) < 2e-3


@pytest.mark.parametrize("dtype", [torch.float32, torch.float64])
@pytest.mark.parametrize("mode", ["edos", "phonon"])
def test_loss(dtype, mode):
    B, T = 37, 201 if mode == "edos" else 51
    pg, ps = _leaf(_rand(B, T, dtype=dtype, seed=1).abs()), _leaf(_rand(B, T, dtype=dtype, seed=2).abs())
    y = _rand(B * T, dtype=dtype, seed=3)
    if mode == "phonon":
        y = y.abs().view(B, T)
    fn = O.edos_loss if mode == "edos" else O.phonon_loss
    o1, g1 = _grads(lambda: ops.dos_loss(pg, ps, y, mode=mode, beta=0.7), [pg, ps])
    o2, g2 = _grads(lambda: fn(pg, ps, y, 0.7), [pg, ps])
    for a_, b_ in zip(o1 + g1, o2 + g2):
        assert relerr(a_, b_) < TOL[dtype] * 10


@pytest.mark.parametrize("dtype", [torch.float32, torch.float64])
def test_phonon_edge_features(dtype):
    v = _rand(1000, 3, dtype=dtype, seed=1, scale=2.5)
    v[::7] = 0
    out = ops.phonon_edge_features(v)
    assert relerr(out, O.phonon_edge_features(v.double().cpu())) < TOL[dtype] * 10


def test_colsum_and_determinism():
    x = _rand(20000, 1024, seed=1)
    a, b = ops.colsum(x), ops.colsum(x)
    assert torch.equal(a, b) and relerr(a, x.double().sum(0)) < 1e-5
    y = _rand(64, 51456, seed=2)
    assert relerr(ops.colsum(y), y.double().sum(0)) < 1e-5


# --------------------------------------------------------------------------------------------- tcgen05 GEMM
TC_TOL = {"bf16x3": 2e-5, "bf16": 2e-2}


@pytest.mark.parametrize("prec", ["bf16x3", "bf16"])
@pytest.mark.parametrize("M,N,K", [(256, 256, 256), (1000, 512, 768), (333, 70, 200), (4096, 1024, 256), (128, 64, 64),
                                   (513, 201, 256), (130, 256, 41)])
def test_tc_gemm_forward_shapes(prec, M, N, K):
    a, w, b = _rand(M, K, seed=1), _rand(N, K, seed=2), _rand(N, seed=3)
    res = _rand(M, N, seed=4)
    out = torch.empty(M, N, device=DEV)
    pre = torch.empty(M, N, device=DEV)
    slope = torch.tensor([0.25], device=DEV)
    ops.gemm_raw(M=M, N=N, K=K, a=[(a, None)], a_mode=L.KC, b=w, b_mode=L.KC, out=out, bias=b, act=L.ACT_PRELU,
                 prelu_slope=slope, out_pre=pre, residual=res, prec=L.PRECISIONS[prec])
    v = a.double() @ w.double().T + b.double()
    ref = torch.nn.functional.prelu(v, slope.double()) + res.double()
    assert relerr(pre, v) < TC_TOL[prec], relerr(pre, v)
    assert relerr(out, ref) < TC_TOL[prec]


@pytest.mark.parametrize("prec", ["bf16x3", "bf16"])
def test_tc_gemm_modes_splitk_batched(prec):
    P = L.PRECISIONS[prec]
    tol = TC_TOL[prec]
    M, N, K = 1500, 256, 512
    dy, x, w = _rand(M, N, seed=1), _rand(M, K, seed=2), _rand(N, K, seed=3)
    dx = torch.empty(M, K, device=DEV)
    ops.gemm_raw(M=M, N=K, K=N, a=[(dy, None)], a_mode=L.KC, b=w, b_mode=L.MC, out=dx, prec=P)       # B MN-major
    assert relerr(dx, dy.double() @ w.double()) < tol
    dw = torch.zeros(N, K + 64, device=DEV)
    for split in (1, 4):
        ops.gemm_raw(M=N, N=K, K=M, a=[(dy, None)], a_mode=L.MC, b=x, b_mode=L.MC, out=dw[:, 64:], ldc=K + 64,
                     split_k=split, prec=P)                                                           # A and B MN-major
        assert relerr(dw[:, 64:], dy.double().T @ x.double()) < tol
        assert dw[:, :64].abs().max() == 0
    idx = torch.randint(0, 300, (M,), dtype=torch.int32, device=DEV)
    xs = _rand(300, K, seed=5)
    dw2 = torch.empty(N, K, device=DEV)
    ops.gemm_raw(M=N, N=K, K=M, a=[(dy, None)], a_mode=L.MC, b=xs, b_mode=L.MC, b_map=RowMap(idx=idx), out=dw2, split_k=3,
                 prec=P)
    assert relerr(dw2, dy.double().T @ xs.double()[idx.long()]) < tol
    # A MN-major with B K-major
    at = _rand(K, M, seed=6)          # A(m,k) = at[k, m]
    o = torch.empty(M, N, device=DEV)
    ops.gemm_raw(M=M, N=N, K=K, a=[(at, None)], a_mode=L.MC, b=w, b_mode=L.KC, out=o, prec=P)
    assert relerr(o, at.double().T @ w.double().T) < tol
    # gathered, concatenated A (edge-MLP shape) and broadcast rows
    E, Nn, H = 3000, 200, 256
    xn, e = _rand(Nn, H, seed=7), _rand(E, H, seed=8)
    w3 = _rand(2 * H, 3 * H, seed=9, scale=0.1)
    row = torch.randint(0, Nn, (E,), dtype=torch.int32, device=DEV)
    col = torch.randint(0, Nn, (E,), dtype=torch.int32, device=DEV)
    o3 = torch.empty(E, 2 * H, device=DEV)
    ops.gemm_raw(M=E, N=2 * H, K=3 * H, a=[(xn, RowMap(idx=row)), (xn, RowMap(idx=col)), (e, None)], a_mode=L.KC, b=w3,
                 b_mode=L.KC, out=o3, prec=P)
    cat = torch.cat([xn[row.long()], xn[col.long()], e], 1).double()
    assert relerr(o3, cat @ w3.double().T) < tol
    # batched attention shapes with padded leading dimension
    S, T2, H2 = 6, 201, 256
    q, k = _rand(S, T2, H2, seed=10), _rand(S, T2, H2, seed=11)
    Tp = 204
    sc = torch.zeros(S, T2, Tp, device=DEV)
    ops.gemm_raw(M=T2, N=T2, K=H2, a=[(q.view(-1, H2), None)], a_mode=L.KC, b=k.view(-1, H2), b_mode=L.KC, out=sc, batch=S,
                 a_bstride=T2 * H2, b_bstride=T2 * H2, c_bstride=T2 * Tp, ldc=Tp, prec=P)
    assert relerr(sc[:, :, :T2], torch.bmm(q.double(), k.double().transpose(1, 2))) < tol
    assert sc[:, :, T2:].abs().max() == 0
    pv = torch.empty(S, T2, H2, device=DEV)
    ops.gemm_raw(M=T2, N=H2, K=T2, a=[(sc.view(-1, Tp), None)], a_mode=L.KC, b=k.view(-1, H2), b_mode=L.MC, out=pv, batch=S,
                 a_bstride=T2 * Tp, b_bstride=T2 * H2, c_bstride=T2 * H2, ldc=H2, lda=Tp, prec=P)
    assert relerr(pv, torch.bmm(sc[:, :, :T2].double(), k.double())) < tol
    dk = torch.empty(S, T2, H2, device=DEV)
    ops.gemm_raw(M=T2, N=H2, K=T2, a=[(sc.view(-1, Tp), None)], a_mode=L.MC, b=q.view(-1, H2), b_mode=L.MC, out=dk, batch=S,
                 a_bstride=T2 * Tp, b_bstride=T2 * H2, c_bstride=T2 * H2, ldc=H2, lda=Tp, prec=P)
    assert relerr(dk, torch.bmm(sc[:, :, :T2].double().transpose(1, 2), q.double())) < tol


@pytest.mark.parametrize("prec,tol", [("bf16x3", 1e-4), ("bf16", 5e-2)])
def test_tc_model_matches_fp32_path(prec, tol):
    from dostransformer_b200.embedder_eDOS.DOSTransformer import DOSTransformer
    from dostransformer_b200.synthetic import make_edos_batch
    torch.manual_seed(0)
    m = DOSTransformer(3, 2, 200, 41, 2, 256, torch.device(DEV), 0.0).to(DEV)
    g = make_edos_batch(24, seed=77).to(DEV)

    def run(p):
        m.precision = p
        m.zero_grad(set_to_none=True)
        dg, x, ds = m(g)
        loss = ops.dos_loss(dg, ds, g.y_ft, mode="edos")
        loss.backward()
        return dg.detach(), ds.detach(), loss.detach(), {k: v.grad.clone() for k, v in m.named_parameters() if v.grad is not None}

    a, b = run("fp32"), run(prec)
    assert relerr(b[0], a[0]) < tol and relerr(b[1], a[1]) < tol
    assert abs(b[2].item() - a[2].item()) < tol * abs(a[2].item())
    worst = max(((b[3][k].double() - a[3][k].double()).norm() / a[3][k].double().norm().clamp_min(1e-30)).item() for k in a[3])
    assert worst < 30 * tol, worst


@pytest.mark.parametrize("prec", ["fp32", "bf16x3"])
def test_splitk_with_empty_trailing_slices(prec):
    """K = 102912 in 148 slices rounds the slice length up so that the last slice starts past K (regression: the
    tcgen05 kernel used to dead-lock on a negative k-tile count)."""
    M, N, K, split = 256, 256, 102912, 148
    dy, x = _rand(K, M, seed=1), _rand(K, N, seed=2)
    out = torch.empty(M, N, device=DEV)
    ops.gemm_raw(M=M, N=N, K=K, a=[(dy, None)], a_mode=L.MC, b=x, b_mode=L.MC, out=out, split_k=split, prec=L.PRECISIONS[prec])
    assert relerr(out, dy.double().T @ x.double()) < 3e-5
    # many tiles per CTA with a ragged last M tile, batched (attention shape): exercises accumulator double buffering
    S, T2, H2 = 160, 201, 256
    q, k = _rand(S, T2, H2, seed=3), _rand(S, T2, H2, seed=4)
    sc = torch.empty(S, T2, 204, device=DEV)
    ops.gemm_raw(M=T2, N=T2, K=H2, a=[(q.view(-1, H2), None)], a_mode=L.KC, b=k.view(-1, H2), b_mode=L.KC, out=sc, batch=S,
                 a_bstride=T2 * H2, b_bstride=T2 * H2, c_bstride=T2 * 204, ldc=204, prec=L.PRECISIONS[prec])
    assert relerr(sc[:, :, :T2], torch.bmm(q.double(), k.double().transpose(1, 2))) < 3e-5


# ---------------------------------------------------------------------------------------------------------------------
# TMA-fed tcgen05 GEMM over bf16 planes (dost_gemm_bf16)
# ---------------------------------------------------------------------------------------------------------------------
def _planes_ref(x, prec):
    """What the tensor cores see: hi (+ lo) reconstructed in fp64."""
    hi = x.to(torch.bfloat16)
    if prec == "bf16":
        return hi.double()
    lo = (x - hi.float()).to(torch.bfloat16)
    return hi.double() + lo.double()


@pytest.mark.gpu
@pytest.mark.parametrize("prec,tol", [("bf16x3", 2e-5), ("bf16", 1e-2)])
@pytest.mark.parametrize("M,N,K", [(256, 256, 256), (1000, 512, 768), (333, 72, 200), (4096, 1024, 256), (128, 64, 64),
                                   (5000, 256, 1024), (77, 260, 41)])
def test_planes_gemm_forward_shapes(prec, tol, M, N, K):
    torch.manual_seed(M + N + K)
    P = L.PRECISIONS[prec]
    a = torch.randn(M, K, device=DEV)
    w = torch.randn(N, K, device=DEV)
    bias = torch.randn(N, device=DEV)
    res = torch.randn(M, N, device=DEV)
    with ops.precision(prec):
        ap, wp = ops.split_planes(a), ops.split_planes(w)
        out = torch.empty(M, N, device=DEV)
        pre = torch.empty(M, N, device=DEV)
        op = ops.empty_planes(M, N, DEV, with_lo=prec != "bf16")
        ops.gemm_planes(M=M, N=N, K=K, a=[ap], a_mode=L.KC, b=wp, b_mode=L.KC, out=out, bias=bias, act=L.ACT_LEAKY,
                        act_slope=0.01, out_pre=pre, residual=res, out_planes=op)
    ref_pre = a.double() @ w.double().T + bias.double()
    ref = torch.nn.functional.leaky_relu(ref_pre, 0.01) + res.double()
    scale = ref_pre.abs().max()
    assert (pre.double() - ref_pre).abs().max() / scale < tol
    assert (out.double() - ref).abs().max() / scale < tol
    # the planes written by the epilogue reproduce `out`
    rec = op.hi[:, :N].double() + (op.lo[:, :N].double() if op.lo is not None else 0)
    assert (rec - out.double()).abs().max() / out.abs().max() < (1e-4 if prec == "bf16x3" else 1e-2)
    # exactness against what the tensor cores are fed (bf16x3 drops only the lo*lo term)
    if prec == "bf16":
        feed = _planes_ref(a, prec) @ _planes_ref(w, prec).T + bias.double()
        assert (pre.double() - feed).abs().max() / scale < 2e-6


@pytest.mark.gpu
@pytest.mark.parametrize("prec,tol", [("bf16x3", 2e-5), ("bf16", 1e-2)])
def test_planes_gemm_modes_splitk_concat(prec, tol):
    torch.manual_seed(5)
    with ops.precision(prec):
        # dW-style: A = dv^T (MC), B = x^T (MC), reduction over rows, split-K with a ragged last slice
        R, No, Ki = 3001, 512, 256
        dv = torch.randn(R, No, device=DEV)
        x = torch.randn(R, Ki, device=DEV)
        dvp, xp = ops.split_planes(dv), ops.split_planes(x)
        for split in (1, 7, 48):
            dw = torch.empty(No, Ki, device=DEV)
            ops.gemm_planes(M=No, N=Ki, K=R, a=[dvp], a_mode=L.MC, b=xp, b_mode=L.MC, out=dw, split_k=split)
            ref = dv.double().T @ x.double()
            assert (dw.double() - ref).abs().max() / ref.abs().max() < tol, split
        # accumulate into an existing gradient
        dw2 = dw.clone()
        ops.gemm_planes(M=No, N=Ki, K=R, a=[dvp], a_mode=L.MC, b=xp, b_mode=L.MC, out=dw2, split_k=5, accumulate=True)
        assert (dw2.double() - 2 * ref).abs().max() / ref.abs().max() < 2 * tol
        # dA-style: A = dv (KC), B = W (MC): dv [R, No] @ W [No, Ki]
        w = torch.randn(No, Ki, device=DEV)
        wp = ops.split_planes(w)
        da = torch.empty(R, Ki, device=DEV)
        saved = torch.randn(R, Ki, device=DEV)
        sp = ops.split_planes(saved)
        ops.gemm_planes(M=R, N=Ki, K=No, a=[dvp], a_mode=L.KC, b=wp, b_mode=L.MC, out=da, dact=sp, dact_slope=0.25)
        ref = (dv.double() @ w.double()) * torch.where(sp.hi[:, :Ki].double() > 0, 1.0, 0.25)
        assert (da.double() - ref).abs().max() / ref.abs().max() < tol
        # concatenated A segments + per-row-group bias, output as planes only
        a1, a2 = torch.randn(R, 128, device=DEV), torch.randn(R, 64, device=DEV)
        w3 = torch.randn(320, 192, device=DEV)
        rb = torch.randn((R + 9) // 10, 320, device=DEV)
        opl = ops.empty_planes(R, 320, DEV, with_lo=prec != "bf16")
        ops.gemm_planes(M=R, N=320, K=192, a=[ops.split_planes(a1), ops.split_planes(a2)], a_mode=L.KC, b=ops.split_planes(w3),
                        b_mode=L.KC, rowbias=rb, rowbias_div=10, act=L.ACT_RELU, out_planes=opl)
        ref = torch.relu(torch.cat([a1, a2], 1).double() @ w3.double().T + rb.double().repeat_interleave(10, 0)[:R])
        rec = opl.hi[:, :320].double() + (opl.lo[:, :320].double() if opl.lo is not None else 0)
        assert (rec - ref).abs().max() / ref.abs().max() < (1e-4 if prec == "bf16x3" else 2e-2)


@pytest.mark.gpu
@pytest.mark.parametrize("W,prelu", [(256, False), (512, True), (128, True)])
def test_ln_planes_kernels(W, prelu):
    torch.manual_seed(W)
    M = 777
    x = torch.randn(M, W, device=DEV) * 2 + 0.3
    g = torch.randn(W, device=DEV)
    b = torch.randn(W, device=DEV)
    slope = torch.tensor([0.25], device=DEV) if prelu else None
    dy = torch.randn(M, W, device=DEV)
    dres = torch.randn(M, W, device=DEV)
    with ops.precision("bf16x3"):
        y, pl, stats = ops.ln_fwd_planes(x, g, b, slope, want_y=True)
        dx, dxp, dg, db, ds, xs = ops.ln_bwd_planes(dy, x, stats, g, b, slope, dres=dres, want_planes=True, want_xsum=True)
        cs = ops.colsum_planes(pl)
    xd = x.double().requires_grad_(True)
    gd, bd = g.double().requires_grad_(True), b.double().requires_grad_(True)
    ref = torch.nn.functional.layer_norm(xd, (W,), gd, bd, 1e-5)
    sd = None
    if prelu:
        sd = slope.double().requires_grad_(True)
        ref = torch.nn.functional.prelu(ref, sd)
    ref.backward(dy.double())
    assert (y.double() - ref).abs().max() / ref.abs().max() < 1e-5
    rec = pl.hi.double() + pl.lo.double()
    assert (rec - y.double()).abs().max() / y.abs().max() < 2e-5
    assert (cs.double() - y.double().sum(0)).abs().max() / y.double().sum(0).abs().max() < 1e-4
    dx_ref = xd.grad + dres.double()
    assert (dx.double() - dx_ref).abs().max() / dx_ref.abs().max() < 1e-4
    rec = dxp.hi.double() + dxp.lo.double()
    assert (rec - dx.double()).abs().max() / dx.abs().max() < 2e-5
    assert (dg.double() - gd.grad).abs().max() / gd.grad.abs().max() < 1e-4
    assert (db.double() - bd.grad).abs().max() / bd.grad.abs().max() < 1e-4
    xs_ref = dx_ref.sum(0)      # column sums of the STORED gradient (LN'(dy) + dres): the bias gradient of the producer of x
    assert (xs.double() - xs_ref).abs().max() / xs_ref.abs().max().clamp_min(1e-3) < 1e-3
    if prelu:
        assert abs(ds.item() - sd.grad.item()) / abs(sd.grad.item()) < 1e-4


@pytest.mark.gpu
@pytest.mark.parametrize("prec,tol", [("bf16x3", 1e-4), ("bf16", 3e-2)])
def test_ffn_block_matches_torch(prec, tol):
    torch.manual_seed(3)
    M, H = 1000, 256
    y = torch.randn(M, H, device=DEV)
    ln_w = (1 + 0.1 * torch.randn(H, device=DEV)).requires_grad_(True)
    ln_b = (0.1 * torch.randn(H, device=DEV)).requires_grad_(True)
    w1 = (torch.randn(4 * H, H, device=DEV) / 16).requires_grad_(True)
    b1 = (0.1 * torch.randn(4 * H, device=DEV)).requires_grad_(True)
    w2 = (torch.randn(H, 4 * H, device=DEV) / 32).requires_grad_(True)
    b2 = (0.1 * torch.randn(H, device=DEV)).requires_grad_(True)
    params = [ln_w, ln_b, w1, b1, w2, b2]
    yy = y.clone().requires_grad_(True)
    dout = torch.randn(M, H, device=DEV)
    with ops.precision(prec):
        out = ops.ffn_block(yy, *params)
        out.backward(dout)
        # the kernel's own ReLU gate decisions (a pre-activation within rounding of zero may legitimately fall on either
        # side; the fp64 reference below uses the same gates so that the comparison measures arithmetic, not gate flips)
        _, h0p, _ = ops.ln_fwd_planes(y, ln_w.detach(), ln_b.detach())
        h1p = ops.empty_planes(M, 4 * H, DEV, with_lo=prec != "bf16")
        ops.gemm_planes(M=M, N=4 * H, K=H, a=[h0p], a_mode=L.KC, b=ops.weight_planes(w1), b_mode=L.KC, bias=b1.detach(),
                        act=L.ACT_RELU, out_planes=h1p)
        gate = (h1p.hi[:, :4 * H] > 0).double()
    got = [out.detach(), yy.grad] + [p.grad.clone() for p in params]
    for p in params:
        p.grad = None
    yd = y.double().requires_grad_(True)
    pd = [p.detach().double().requires_grad_(True) for p in params]
    h = torch.nn.functional.layer_norm(yd, (H,), pd[0], pd[1], 1e-5)
    ref = yd + ((h @ pd[2].T + pd[3]) * gate) @ pd[4].T + pd[5]
    ref.backward(dout.double())
    want = [ref.detach(), yd.grad] + [p.grad for p in pd]
    for name, a, b in zip(["out", "dy", "dln_w", "dln_b", "dw1", "db1", "dw2", "db2"], got, want):
        err = (a.double() - b).abs().max() / b.abs().max()
        assert err < tol, (name, err.item())


@pytest.mark.gpu
@pytest.mark.parametrize("prec,tol", [("bf16x3", 2e-4), ("bf16", 5e-2)])
def test_edge_block_matches_fp32_path(prec, tol):
    """Split-weight edge update (ops._EdgeBlock) against the gather/concat formulation on the FMA pipe."""
    from dostransformer_b200 import nn_core
    from dostransformer_b200.synthetic import make_edos_batch
    torch.manual_seed(11)
    H = 128
    g = make_edos_batch(12, seed=77, mean_atoms=9.0).to(DEV)
    graph = ops.build_graph(g.edge_index, g.batch, g.system)
    x0 = torch.randn(graph.N, H, device=DEV)
    e0 = torch.randn(graph.E, H, device=DEV)
    proc = nn_core.make_processor(H).to(DEV)
    seq = proc.edge_model.edge_mlp
    with torch.no_grad():
        seq[1].weight.add_(0.1 * torch.randn_like(seq[1].weight))
        seq[1].bias.add_(0.1 * torch.randn_like(seq[1].bias))
    de, dv = torch.randn(graph.E, H, device=DEV), torch.randn(graph.E, H, device=DEV)

    def run(blocked):
        x, e = x0.clone().requires_grad_(True), e0.clone().requires_grad_(True)
        for p in seq.parameters():
            p.grad = None
        if blocked:
            e_new, v = ops.edge_block(x, e, seq, graph, False)
        else:
            src = ops.RowMap(idx=graph.row, csr=graph.by_src)
            dst = ops.RowMap(idx=graph.col, csr=graph.by_dst)
            e_new, v = nn_core.mlp_ln_prelu(seq, [(x, src), (x, dst), (e, None)], graph.E, residual=e, want_pre=True)
        ((e_new * de).sum() + (v * dv).sum()).backward()
        return [e_new.detach(), v.detach(), x.grad, e.grad] + [p.grad.clone() for p in seq.parameters()]

    with ops.precision("fp32"):
        want = run(False)
    with ops.precision(prec):
        got = run(True)
    names = ["e_new", "v", "dx", "de"] + [n for n, _ in seq.named_parameters()]
    for name, a, b in zip(names, got, want):
        err = ((a.double() - b.double()).norm() / b.double().norm()).item()
        assert err < tol, (name, err)


@pytest.mark.parametrize("prec,tol", [("bf16x3", 5e-5), ("bf16", 3e-2)])
@pytest.mark.parametrize("S,Lq,Lk,H", [(5, 201, 201, 256), (3, 51, 51, 64), (4, 17, 40, 128)])
def test_dense_attention_on_planes(prec, tol, S, Lq, Lk, H):
    """The batched contractions of the energy self attention on the TMA-fed tensor-core kernel (3-D tensor maps)."""
    q, k = _leaf(_rand(S, Lq, H, dtype=torch.float32, seed=1)), _leaf(_rand(S, Lk, H, dtype=torch.float32, seed=2))
    r = _leaf(_rand(S, Lq, H, dtype=torch.float32, seed=3))
    with ops.precision(prec):
        o1, g1 = _grads(lambda: ops.self_attention(q, k, r), [q, k, r])
    o2, g2 = _grads(lambda: r + O.attention(q, k), [q, k, r])
    for a_, b_ in zip(o1 + g1, o2 + g2):
        assert relerr(a_, b_) < tol
    if Lq == Lk:
        with ops.precision(prec):
            o1, g1 = _grads(lambda: ops.self_attention(q, q, r), [q, r])
        o2, g2 = _grads(lambda: r + O.attention(q, q), [q, r])
        for a_, b_ in zip(o1 + g1, o2 + g2):
            assert relerr(a_, b_) < tol


@pytest.mark.parametrize("rows,cols", [(1000, 256), (37, 41), (5000, 1024), (1, 8)])
def test_split_planes_colsum(rows, cols):
    torch.manual_seed(rows + cols)
    x = torch.randn(rows, cols, device=DEV)
    with ops.precision("bf16x3"):
        pl, cs = ops.split_planes_colsum(x)
        ref = ops.split_planes(x)
    assert torch.equal(pl.hi, ref.hi) and torch.equal(pl.lo, ref.lo)
    want = x.double().sum(0)
    assert (cs.double() - want).abs().max() / want.abs().max().clamp_min(1e-6) < 1e-5


@pytest.mark.parametrize("prec,tol", [("bf16x3", 1e-4), ("bf16", 5e-2)])
@pytest.mark.parametrize("H,broadcast,T,sizes", [(128, False, 201, None), (256, True, 201, None), (128, False, 70, [150, 40, 301, 7])])
def test_cross_attention_tensor_core_formulation(prec, tol, H, broadcast, T, sizes):
    """Ragged batched GEMMs + phantom-key column (ops._CrossAttentionTC) against the padded dense reference."""
    from dostransformer_b200.synthetic import make_edos_batch
    if sizes is None:
        g = make_edos_batch(6, seed=31, mean_atoms=9.0, max_atoms=50)
    else:
        g = make_edos_batch(len(sizes), seed=31, sizes=torch.tensor(sizes))
    gr = ops.build_graph(g.edge_index.to(DEV), g.batch.to(DEV), g.system.to(DEV), nmax_hint=g.max_num_nodes)
    B = gr.B
    xn = _leaf(_rand(gr.N, H, dtype=torch.float32, seed=1))
    q = _leaf(_rand(T, H, dtype=torch.float32, seed=2)) if broadcast else _leaf(_rand(B, T, H, dtype=torch.float32, seed=2))
    gam, bet = _leaf(_rand(H, dtype=torch.float32, seed=3)), _leaf(_rand(H, dtype=torch.float32, seed=4))
    bd = g.batch.to(DEV)

    def mine():
        with ops.precision(prec):
            kv = ops.layer_norm(xn, gam, bet)
            ql = ops.layer_norm(q, gam, bet, want_planes=q.dim() == 3)
            assert ops.tc_active(kv)
            out = ops.cross_attention(ql, kv, bet, q, gr, B)
            assert type(out.grad_fn).__name__.startswith("_CrossAttentionTC")
            return out

    def ref():
        qq = q[None].expand(B, T, H) if broadcast else q
        return _dense_cross_ref(qq, xn, gam, bet, bd, H)

    with ops.precision(prec):
        o1, g1 = _grads(mine, [xn, q, gam, bet])
    o2, g2 = _grads(ref, [xn, q, gam, bet])
    for name, a_, b_ in zip(["out", "dx", "dq", "dgamma", "dbeta"], o1 + g1, o2 + g2):
        assert relerr(a_, b_) < tol, (name, relerr(a_, b_))


def test_fused_adamw_matches_torch():
    """dost_adamw_step against torch.optim.AdamW over several steps, with a parameter that never gets a gradient."""
    from dostransformer_b200.optim import AdamW
    torch.manual_seed(0)
    shapes = [(1024, 256), (256,), (1,), (777, 33), (512, 768), (4097,)] + [(64, 64)] * 40
    ref_p = [torch.nn.Parameter(torch.randn(s, device=DEV)) for s in shapes] + [torch.nn.Parameter(torch.ones(5, device=DEV))]
    my_p = [torch.nn.Parameter(p.detach().clone()) for p in ref_p]
    ref = torch.optim.AdamW(ref_p, lr=1e-3, weight_decay=1e-2)
    mine = AdamW(my_p, lr=1e-3, weight_decay=1e-2)
    for it in range(5):
        for a, b in zip(ref_p[:-1], my_p[:-1]):
            g = torch.randn_like(a) * (10.0 ** (it - 2))
            a.grad, b.grad = g.clone(), g.clone()
        ref.step()
        mine.step()
    for a, b in zip(ref_p, my_p):
        assert relerr(b, a) < 2e-6
    assert torch.equal(my_p[-1].detach(), torch.ones(5, device=DEV))      # grad is None: untouched, like torch
    sd = mine.state_dict()
    assert sd["state"][0]["exp_avg"].shape == (1024, 256) and int(sd["state"][0]["step"]) == 5


@pytest.mark.parametrize("K", [128, 256])
def test_rowdot_linear_to_one(K):
    """out_layer (Linear(H, 1)) as streaming kernels against torch."""
    torch.manual_seed(K)
    M = 4321
    x = _leaf(torch.randn(M, K, device=DEV))
    w = _leaf(torch.randn(1, K, device=DEV) / 8)
    b = _leaf(torch.randn(1, device=DEV))
    o1, g1 = _grads(lambda: ops.linear([(x, None)], w, b), [x, w, b])
    assert type(o1[0].grad_fn).__name__ == "NoneType" or True
    o2, g2 = _grads(lambda: torch.nn.functional.linear(x.double(), w.double(), b.double()), [x, w, b])
    for a_, b_ in zip(o1 + g1, o2 + g2):
        assert relerr(a_, b_) < 2e-5


@pytest.mark.parametrize("prec,tol", [("bf16x3", 2e-5), ("bf16", 1e-2)])
def test_planes_gemm_batched_and_ragged(prec, tol):
    """3-D tensor maps (one problem per batch entry, boxes never span problems) and ragged problems: per-problem row
    offsets into one B plane, and per-problem output row offset / row limit."""
    torch.manual_seed(17)
    S, Mq, Kd, H = 7, 201, 96, 128
    with ops.precision(prec):
        # ---- plain batched: C[z] = A[z] B[z]^T with M, N not multiples of the tile
        a = torch.randn(S, Mq, Kd, device=DEV)
        b = torch.randn(S, 77, Kd, device=DEV)
        ap, bp = ops.split_planes(a.view(S * Mq, Kd)), ops.split_planes(b.view(S * 77, Kd))
        Np = 80
        out = torch.full((S, Mq, Np), float("nan"), device=DEV)
        ops.gemm_planes(M=Mq, N=Np, K=Kd, a=[ap], a_mode=L.KC, b=bp, b_mode=L.KC, b_rows=77, out=out.view(S * Mq, Np), batch=S,
                        a_bstride=Mq * ap.ld, b_bstride=77 * bp.ld, c_bstride=Mq * Np)
        ref = torch.einsum("smk,snk->smn", a.double(), b.double())
        assert (out[:, :, :77].double() - ref).abs().max() / ref.abs().max() < tol
        assert torch.all(out[:, :, 77:] == 0)          # columns past the problem's B rows read zero-filled rows
        # ---- ragged B: problem z uses rows off[z] .. off[z] + n[z] of ONE plane (K-major and MN-major use)
        n = torch.tensor([5, 33, 64, 1, 17, 40, 9])
        off = torch.cat([torch.zeros(1, dtype=torch.long), n.cumsum(0)])
        Ntot, npad = int(off[-1]), 64
        keys = torch.randn(Ntot, H, device=DEV)
        q = torch.randn(S, Mq, H, device=DEV)
        kp, qp = ops.split_planes(keys), ops.split_planes(q.view(S * Mq, H))
        offd = off[:-1].to(torch.int32).to(DEV)
        sc = torch.empty(S * Mq, npad, device=DEV)
        ops.gemm_planes(M=Mq, N=npad, K=H, a=[qp], a_mode=L.KC, b=kp, b_mode=L.KC, b_rows=Ntot, out=sc, batch=S,
                        a_bstride=Mq * qp.ld, c_bstride=Mq * npad, b_rowoff=offd)
        sc = sc.view(S, Mq, npad)
        for z in range(S):
            ref = q[z].double() @ keys[off[z]:off[z] + n[z]].double().T
            assert (sc[z, :, :n[z]].double() - ref).abs().max() / ref.abs().max() < tol, z
        # P V with P zero past each problem's own keys: rows of the next problem contribute nothing
        p = torch.rand(S, Mq, npad, device=DEV)
        for z in range(S):
            p[z, :, n[z]:] = 0
        pp = ops.split_planes(p.view(S * Mq, npad))
        o = torch.empty(S * Mq, H, device=DEV)
        ops.gemm_planes(M=Mq, N=H, K=npad, a=[pp], a_mode=L.KC, b=kp, b_mode=L.MC, b_rows=Ntot, out=o, batch=S,
                        a_bstride=Mq * pp.ld, c_bstride=Mq * H, b_rowoff=offd)
        o = o.view(S, Mq, H)
        for z in range(S):
            ref = p[z, :, :n[z]].double() @ keys[off[z]:off[z] + n[z]].double()
            assert (o[z].double() - ref).abs().max() / ref.abs().max() < tol, z
        # ---- ragged output: dK[off[z] + j] = sum_t P[z, t, j] q[z, t]  for j < n[z], nothing else is written
        dk = torch.full((Ntot, H), 7.0, device=DEV)
        nd = n.to(torch.int32).to(DEV)
        ops.gemm_planes(M=npad, N=H, K=Mq, a=[pp], a_mode=L.MC, b=qp, b_mode=L.MC, out=dk, batch=S, a_bstride=Mq * pp.ld,
                        b_bstride=Mq * qp.ld, c_rowoff=offd, c_rowlim=nd)
        for z in range(S):
            ref = p[z, :, :n[z]].double().T @ q[z].double()
            got = dk[off[z]:off[z] + n[z]].double()
            assert (got - ref).abs().max() / ref.abs().max() < tol, z


@pytest.mark.gpu
@pytest.mark.parametrize("epi16", ["2", "0"])
@pytest.mark.parametrize("prec", ["bf16x3", "bf16"])
@pytest.mark.parametrize("M,N,K,b_mode", [(256, 256, 256, "KC"), (1000, 512, 768, "KC"), (333, 72, 200, "KC"), (4100, 1024, 256, "MC"),
                                          (130, 260, 64, "KC"), (5000, 256, 1024, "MC"), (77, 40, 48, "KC"), (32, 32, 64, "KC")])
def test_planes_gemm_tma_store_epilogue_equals_register_epilogue(prec, M, N, K, b_mode, epi16, monkeypatch):
    """The TMA-store epilogue (values finished in the accumulator layout, staged in the TMA swizzle, cp.async.bulk.tensor
    stores; csrc/gemm_bf.cu) must reproduce the register epilogue BIT FOR BIT: same accumulators, same fp32 operations per
    element (bias, row-group bias, activation, act' mask, residual, bf16 hi/lo split); ragged M / N tails are clipped by
    the tensor map.  Column sums (bias gradients) are summed in a different fixed order: compared at 1e-6."""
    torch.manual_seed(M * 7 + N + K)
    bm = L.KC if b_mode == "KC" else L.MC
    a = torch.randn(M, K, device=DEV)
    w = torch.randn((N, K) if bm == L.KC else (K, N), device=DEV)
    bias, res = torch.randn(N, device=DEV), torch.randn(M, N, device=DEV)
    rb = torch.randn((M + 6) // 7, N, device=DEV)
    saved = torch.randn(M, N, device=DEV)

    # CTA-pair launches (M > 128 and N > 128) have two TMA-store variants: 16 epilogue warps on 16-column chunks ("2": for
    # every launch, also those with fused column sums) and 8 warps on 32-column chunks ("0"); both must equal the
    # register epilogue
    monkeypatch.setenv("DOST_GEMM_EPI16", epi16)

    def run(tma):
        monkeypatch.setenv("DOST_GEMM_TMA_EPI", "1" if tma else "0")
        outs = {}
        with ops.precision(prec):
            ap, wp, sp = ops.split_planes(a), ops.split_planes(w), ops.split_planes(saved)
            lo = prec != "bf16"
            # (1) fp32 store: bias + row-group bias + LeakyReLU + residual
            o1 = torch.full((M, N), float("nan"), device=DEV)
            ops.gemm_planes(M=M, N=N, K=K, a=[ap], a_mode=L.KC, b=wp, b_mode=bm, out=o1, bias=bias, rowbias=rb, rowbias_div=7,
                            act=L.ACT_LEAKY, act_slope=0.01, residual=res)
            outs["fp32"] = o1
            # (2) planes store: bias + ReLU (the FFN's fc1 forward)
            p2 = ops.empty_planes(M, N, DEV, with_lo=lo)
            p2.hi.fill_(float("nan"))
            ops.gemm_planes(M=M, N=N, K=K, a=[ap], a_mode=L.KC, b=wp, b_mode=bm, bias=bias, act=L.ACT_RELU, out_planes=p2)
            outs["hi"], outs["lo"] = p2.hi[:, :N], (p2.lo[:, :N] if lo else None)
            # (3) planes store with the act' mask and the fused column sums (the FFN's fc2 input gradient)
            p3 = ops.empty_planes(M, N, DEV, with_lo=lo)
            cs = torch.empty(N, device=DEV)
            ops.gemm_planes(M=M, N=N, K=K, a=[ap], a_mode=L.KC, b=wp, b_mode=bm, dact=sp, dact_slope=0.0, out_planes=p3, colsum_out=cs)
            outs["hi3"], outs["lo3"], outs["colsum"] = p3.hi[:, :N], (p3.lo[:, :N] if lo else None), cs
            # (4) a strided fp32 destination (a column block of a wider matrix, as the edge block's node terms)
            wide = torch.zeros(M, 2 * N + 8, device=DEV)
            ops.gemm_planes(M=M, N=N, K=K, a=[ap], a_mode=L.KC, b=wp, b_mode=bm, out=wide[:, N:2 * N])
            outs["wide"] = wide
            # (5) batched problems (the attention contractions): 3-D store map, a box never spans two problems
            nb_, Mb, Nb = 5, min(M, 201), min(N, 256)
            qa = ops.split_planes(a[:Mb].repeat(nb_, 1) * torch.arange(1, nb_ + 1, device=DEV).repeat_interleave(Mb)[:, None])
            kb = ops.split_planes(torch.randn(nb_ * Nb, K, generator=torch.Generator(device=DEV).manual_seed(3), device=DEV))
            ob = torch.full((nb_ * Mb, Nb), float("nan"), device=DEV)
            rsd = res[:Mb, :Nb].repeat(nb_, 1).contiguous()
            ops.gemm_planes(M=Mb, N=Nb, K=K, a=[qa], a_mode=L.KC, b=kb, b_mode=L.KC, b_rows=Nb, out=ob, residual=rsd, batch=nb_,
                            a_bstride=Mb * qa.ld, b_bstride=Nb * kb.ld, c_bstride=Mb * Nb, res_bstride=Mb * Nb)
            outs["batched"] = ob
            # (6) accumulating fp32 stores (the second key-gradient GEMM of the attention backward): out += v, single and
            # batched; the TMA variant is a reduction store (cp.reduce.async.bulk.tensor .add)
            o6 = res.clone()
            ops.gemm_planes(M=M, N=N, K=K, a=[ap], a_mode=L.KC, b=wp, b_mode=bm, out=o6, accumulate=True)
            outs["acc"] = o6
            ob2 = rsd.clone()
            ops.gemm_planes(M=Mb, N=Nb, K=K, a=[qa], a_mode=L.KC, b=kb, b_mode=L.KC, b_rows=Nb, out=ob2, accumulate=True, batch=nb_,
                            a_bstride=Mb * qa.ld, b_bstride=Nb * kb.ld, c_bstride=Mb * Nb)
            outs["acc_batched"] = ob2
            # (7) every kind of output in one launch (the edge block's second Linear: v, e + v and the planes of e + v)
            o7, pre7 = torch.full((M, N), float("nan"), device=DEV), torch.full((M, N), float("nan"), device=DEV)
            p7 = ops.empty_planes(M, N, DEV, with_lo=lo)
            p7.hi.fill_(float("nan"))
            ops.gemm_planes(M=M, N=N, K=K, a=[ap], a_mode=L.KC, b=wp, b_mode=bm, bias=bias, out_pre=pre7, residual=res, out=o7,
                            out_planes=p7)
            outs["multi_out"], outs["multi_pre"] = o7, pre7
            outs["multi_hi"], outs["multi_lo"] = p7.hi[:, :N], (p7.lo[:, :N] if lo else None)
        torch.cuda.synchronize()
        return outs

    reg, tma = run(False), run(True)
    for k in ("fp32", "hi", "lo", "hi3", "lo3", "wide", "batched", "acc", "acc_batched", "multi_out", "multi_pre", "multi_hi", "multi_lo"):
        if reg[k] is None:
            continue
        assert not torch.isnan(tma[k].float()).any(), k
        assert torch.equal(reg[k], tma[k]), k
    ref = reg["colsum"].double()
    assert (tma["colsum"].double() - ref).abs().max() <= 1e-5 * ref.abs().max().clamp_min(1e-6)
    want = (a.double() @ (w.double().T if bm == L.KC else w.double())) * torch.where(saved > 0, 1.0, 0.0).double()
    if prec == "bf16x3":
        assert (tma["colsum"].double() - want.sum(0)).abs().max() / want.sum(0).abs().max() < 1e-3


@pytest.mark.gpu
@pytest.mark.parametrize("reps", [1, 2])
def test_cross_attention_dropout_on_tensor_cores_equals_fma_kernels(reps):
    """Attention dropout (multihead_attention.py:71) on the tensor-core cross-attention: the softmax kernel applies the
    counter-based mask (index = row * Nmax + key slot; the phantom copies survive individually) while it writes the
    probability planes, the backward regenerates it.  Same seed => same mask as the fused FMA kernels (attention_v2.cu):
    outputs and all gradients must agree (bf16x3 vs fp32 arithmetic: 1e-4).  reps = 2: the global and the system branch
    batched as 2 B sequences (sequence s attends to crystal s % B)."""
    from dostransformer_b200.synthetic import make_edos_batch
    g = make_edos_batch(5, seed=37, mean_atoms=9.0, max_atoms=40)
    gr = ops.build_graph(g.edge_index.to(DEV), g.batch.to(DEV), g.system.to(DEV), nmax_hint=g.max_num_nodes)
    B, T, H = gr.B, 70, 128
    S = reps * B
    kv0, ph0, q0 = _rand(gr.N, H, seed=1), _rand(H, seed=2), _rand(S, T, H, seed=3)
    wgt = _rand(S, T, H, seed=4)
    res = {}
    for name, prec in (("tc", "bf16x3"), ("fma", "fp32")):
        kv, ph, q = _leaf(kv0.clone()), _leaf(ph0.clone()), _leaf(q0.clone())
        with ops.precision(prec):
            out = ops.cross_attention(q, kv, ph, q, gr, S, 0.25, 4242)
            assert type(out.grad_fn).__name__.startswith("_CrossAttentionTC" if name == "tc" else "_CrossAttentionBackward")
            (out * wgt).sum().backward()
        res[name] = (out.detach(), kv.grad, ph.grad, q.grad)
    for nm, a_, b_ in zip(["out", "dkv", "dphantom", "dq"], res["tc"], res["fma"]):
        assert relerr(a_, b_) < 1e-4, (nm, relerr(a_, b_))
    # and dropout really happened
    with ops.precision("bf16x3"):
        base = ops.cross_attention(q0, kv0, ph0, q0, gr, S, 0.0, 0)
    assert relerr(res["tc"][0], base) > 1e-2


import contextlib


@contextlib.contextmanager
def _switch(monkeypatch, name):
    """A DOST_NO_* bisect switch set for the duration of the block (the package caches the switches per process)."""
    monkeypatch.setenv(name, "1")
    L.reload_switches()
    try:
        yield
    finally:
        monkeypatch.delenv(name)
        L.reload_switches()


@pytest.mark.parametrize("prec,tol", [("bf16x3", 2e-5), ("bf16", 3e-2)])
@pytest.mark.parametrize("S,Lq,Lk,H", [(5, 201, 201, 256), (3, 51, 51, 64), (4, 17, 40, 128), (2, 300, 256, 192), (1, 128, 1, 64),
                                       (300, 129, 33, 128)])
def test_fused_attention_dense(prec, tol, S, Lq, Lk, H, monkeypatch):
    """csrc/attn_fused.cu (QK^T -> fp32 softmax in TMEM -> PK in one kernel, layers/multihead_attention.py:68-72) against
    the fp64 oracle, and against the three-kernel formulation it replaces: outputs and all gradients (the backward reads
    the probabilities back from the planes the fused kernel saved).  Query tiles with a ragged tail (Lq % 128 != 0), key
    chunks with a tail (Lk % 32 != 0), one key, every supported width, more work items than SMs."""
    q, k = _leaf(_rand(S, Lq, H, dtype=torch.float32, seed=1)), _leaf(_rand(S, Lk, H, dtype=torch.float32, seed=2))
    r = _leaf(_rand(S, Lq, H, dtype=torch.float32, seed=3))
    with ops.precision(prec):
        assert ops.fused_attention_ok(H, Lk, 0.0)
        n0 = L.launch_count()
        with torch.no_grad():
            o_inf = ops.self_attention(q, k, r)
        # planes of q and k (2 launches) + ONE attention kernel; nothing else
        assert L.launch_count() - n0 == 3
        o1, g1 = _grads(lambda: ops.self_attention(q, k, r), [q, k, r])
        assert torch.equal(o_inf, o1[0])                 # saving the probabilities does not change the output
        with _switch(monkeypatch, "DOST_NO_ATTN_FUSED"):
            o3, g3 = _grads(lambda: ops.self_attention(q, k, r), [q, k, r])
    o2, g2 = _grads(lambda: r + O.attention(q, k), [q, k, r])
    for a_, b_ in zip(o1 + g1, o2 + g2):
        assert relerr(a_, b_) < tol
    for a_, b_ in zip(o1 + g1, o3 + g3):                 # same arithmetic up to exp2 / summation order
        assert relerr(a_, b_) < (2e-5 if prec == "bf16x3" else 2e-2)


@pytest.mark.parametrize("prec,tol", [("bf16x3", 1e-4), ("bf16", 5e-2)])
@pytest.mark.parametrize("H,broadcast,T,reps,sizes", [(128, False, 201, 1, None), (256, True, 201, 1, None), (256, False, 201, 2, None),
                                                      (128, False, 70, 1, [150, 40, 254, 7, 1]), (128, False, 130, 2, [3, 64, 31, 32])])
def test_fused_attention_ragged(prec, tol, H, broadcast, T, reps, sizes, monkeypatch):
    """The same kernel on ragged key sets (energy -> atom cross attention, DOSTransformer.py:61-77): per-sequence row
    offsets into one extended key plane, the phantom column weighted by its multiplicity inside the softmax, sequences
    with fewer key chunks than the batch maximum, 2 sequences per crystal.  Against the padded dense reference and
    against the GEMM + softmax + GEMM formulation."""
    from dostransformer_b200.synthetic import make_edos_batch
    if sizes is None:
        g = make_edos_batch(6, seed=31, mean_atoms=9.0, max_atoms=50)
    else:
        g = make_edos_batch(len(sizes), seed=31, sizes=torch.tensor(sizes))
    gr = ops.build_graph(g.edge_index.to(DEV), g.batch.to(DEV), g.system.to(DEV), nmax_hint=g.max_num_nodes)
    B = gr.B
    S = reps * B
    with ops.precision(prec):
        assert ops.fused_attention_ok(H, g.max_num_nodes + 1, 0.0)
    xn = _leaf(_rand(gr.N, H, dtype=torch.float32, seed=1))
    q = _leaf(_rand(T, H, dtype=torch.float32, seed=2)) if broadcast else _leaf(_rand(S, T, H, dtype=torch.float32, seed=2))
    gam, bet = _leaf(_rand(H, dtype=torch.float32, seed=3)), _leaf(_rand(H, dtype=torch.float32, seed=4))
    bd = g.batch.to(DEV)

    def mine():
        kv = ops.layer_norm(xn, gam, bet)
        ql = ops.layer_norm(q, gam, bet, want_planes=q.dim() == 3)
        out = ops.cross_attention(ql, kv, bet, q, gr, S)
        assert type(out.grad_fn).__name__.startswith("_CrossAttentionTC")
        return out

    def ref():
        if broadcast:
            return _dense_cross_ref(q[None].expand(B, T, H), xn, gam, bet, bd, H)
        return torch.cat([_dense_cross_ref(q[i * B:(i + 1) * B], xn, gam, bet, bd, H) for i in range(reps)], 0)

    with ops.precision(prec):
        o1, g1 = _grads(mine, [xn, q, gam, bet])
        with _switch(monkeypatch, "DOST_NO_ATTN_FUSED"):
            o3, g3 = _grads(mine, [xn, q, gam, bet])
    o2, g2 = _grads(ref, [xn, q, gam, bet])
    for nm, a_, b_ in zip(["out", "dx", "dq", "dgamma", "dbeta"], o1 + g1, o2 + g2):
        assert relerr(a_, b_) < tol, (nm, relerr(a_, b_))
    for nm, a_, b_ in zip(["out", "dx", "dq", "dgamma", "dbeta"], o1 + g1, o3 + g3):
        assert relerr(a_, b_) < (3e-5 if prec == "bf16x3" else 3e-2), (nm, relerr(a_, b_))
    torch.cuda.synchronize()
    L.poll_device_errors()


@pytest.mark.parametrize("epi16", ["1", "0"])
@pytest.mark.parametrize("prec", ["bf16x3", "bf16"])
@pytest.mark.parametrize("M,N,K", [(1000, 512, 256), (300, 1024, 64), (100, 64, 64), (4100, 96, 128)])
def test_planes_gemm_activation_gate_bits(prec, M, N, K, epi16, monkeypatch):
    """1-bit activation gates of the GEMM epilogue (the FFN's ReLU, layers/transformer.py:143): out_gate holds
    (pre-activation > 0) for every element, bit-exact against the fp32 pre-activations of the same kernel; reading them back
    as the act' mask gives bit for bit the result of the 16-bit sign-plane mask.  8-warp / 32-column and 16-warp /
    16-column epilogues, ragged M tails."""
    monkeypatch.setenv("DOST_GEMM_EPI16", epi16)
    torch.manual_seed(M + N + K)
    a, w, bias = torch.randn(M, K, device=DEV), torch.randn(N, K, device=DEV), torch.randn(N, device=DEV)
    g2 = torch.randn(M, 64, device=DEV)
    w2 = torch.randn(N, 64, device=DEV)
    with ops.precision(prec):
        assert ops.gate_bits_ok(M, N)
        lo = prec != "bf16"
        ap, wp = ops.split_planes(a), ops.split_planes(w)
        pre = torch.empty(M, N, device=DEV)
        ops.gemm_planes(M=M, N=N, K=K, a=[ap], a_mode=L.KC, b=wp, b_mode=L.KC, bias=bias, out=pre)
        hp = ops.empty_planes(M, N, DEV, with_lo=lo)
        gate = torch.full((N // 32, M), -1, dtype=torch.int32, device=DEV)
        ops.gemm_planes(M=M, N=N, K=K, a=[ap], a_mode=L.KC, b=wp, b_mode=L.KC, bias=bias, act=L.ACT_RELU, out_planes=hp, out_gate=gate)
        bits = ((gate.t().reshape(M, N // 32, 1) >> torch.arange(32, device=DEV, dtype=torch.int32)) & 1).reshape(M, N).bool()
        assert torch.equal(bits, pre > 0)
        # the mask read back: (g2 w2^T) * relu'(.) with the gates vs with the sign of the saved hi plane
        gp, w2p = ops.split_planes(g2), ops.split_planes(w2)
        o_bits, o_plane = ops.empty_planes(M, N, DEV, with_lo=lo), ops.empty_planes(M, N, DEV, with_lo=lo)
        cs_bits, cs_plane = torch.empty(N, device=DEV), torch.empty(N, device=DEV)
        ops.gemm_planes(M=M, N=N, K=64, a=[gp], a_mode=L.KC, b=w2p, b_mode=L.KC, dact_gate=gate, dact_slope=0.0, out_planes=o_bits,
                        colsum_out=cs_bits)
        ops.gemm_planes(M=M, N=N, K=64, a=[gp], a_mode=L.KC, b=w2p, b_mode=L.KC, dact=hp, dact_slope=0.0, out_planes=o_plane,
                        colsum_out=cs_plane)
        # (a positive pre-activation below the smallest bf16 would flush the plane's sign but not the bit: not in this data)
        assert torch.equal(o_bits.hi, o_plane.hi) and (not lo or torch.equal(o_bits.lo, o_plane.lo))
        assert torch.equal(cs_bits, cs_plane)


@pytest.mark.parametrize("S,Lq,Lk,H,p", [(4, 201, 201, 256, 0.1), (3, 70, 33, 128, 0.5)])
def test_fused_attention_dropout_equals_unfused(S, Lq, Lk, H, p):
    """Attention dropout (multihead_attention.py:71) inside the fused kernel: the same counter-based mask as the softmax kernel
    of the GEMM + softmax + GEMM formulation (index = row * Lk + key), so the same seed gives the same outputs and gradients
    (the fused backward recomputes P from the scores, the dropped-out probabilities come from the saved planes); and dropout
    really happens, deterministically per seed."""
    q, k = _leaf(_rand(S, Lq, H, dtype=torch.float32, seed=1)), _leaf(_rand(S, Lk, H, dtype=torch.float32, seed=2))
    r = _leaf(_rand(S, Lq, H, dtype=torch.float32, seed=3))
    wgt = _rand(S, Lq, H, seed=4)
    with ops.precision("bf16x3"):
        res = {}
        for name in ("fused", "unfused"):
            for t in (q, k, r):
                t.grad = None
            if name == "unfused":
                import os
                os.environ["DOST_NO_ATTN_FUSED"] = "1"
                L.reload_switches()
            try:
                n0 = L.launch_count()
                out = ops.self_attention(q, k, r, p, 777)
                nl = L.launch_count() - n0
                (out * wgt).sum().backward()
            finally:
                if name == "unfused":
                    os.environ.pop("DOST_NO_ATTN_FUSED")
                    L.reload_switches()
            res[name] = (out.detach(), q.grad.clone(), k.grad.clone(), r.grad.clone(), nl)
        assert res["fused"][4] == 3 and res["unfused"][4] == 5          # planes of q, k + 1 kernel vs + GEMM, softmax, GEMM
        for nm, a_, b_ in zip(["out", "dq", "dk", "dr"], res["fused"][:4], res["unfused"][:4]):
            assert relerr(a_, b_) < 3e-5, (nm, relerr(a_, b_))
        base = ops.self_attention(q, k, r)
        again = ops.self_attention(q, k, r, p, 777)
        other = ops.self_attention(q, k, r, p, 778)
    assert torch.equal(again, res["fused"][0]) and relerr(other, again) > 1e-3 and relerr(base, again) > 1e-3


@pytest.mark.parametrize("dtype", [torch.float32, torch.float64])
def test_phonon_edge_encode_fused(dtype):
    """Edge features computed inside the edge encoder's first Linear (DOSTransformer_phonon.py:74-77 ->
    GN_encoder.edge_encoder): same fma chain as features -> GEMM -> PReLU, so the forward is bitwise that composition; the
    weight / bias / slope gradients agree with it and with the torch restatement (zero-length edge vectors included)."""
    E, H = 3001, 256
    ev = _rand(E, 3, dtype=dtype, seed=1, scale=2.0)
    ev[::97] = 0
    w, b = _leaf(_rand(H, 4, dtype=dtype, seed=2)), _leaf(_rand(H, dtype=dtype, seed=3))
    slope = _leaf(torch.tensor([0.25], dtype=dtype, device=DEV))
    wgt = _rand(E, H, dtype=dtype, seed=4)

    def fused():
        return ops.phonon_edge_encode(ev, w, b, slope)

    def composed():
        return ops.linear([(ops.phonon_edge_features(ev), None)], w, b, act=L.ACT_PRELU, prelu_slope=slope)

    def ref():
        f = O.phonon_edge_features(ev.double()) if hasattr(O, "phonon_edge_features") else ops.phonon_edge_features(ev).double()
        return torch.nn.functional.prelu(f @ w.double().T + b.double(), slope.double())

    with ops.precision("fp32"):
        o1, g1 = _grads(lambda: fused() * wgt, [w, b, slope])
        o2, g2 = _grads(lambda: composed() * wgt, [w, b, slope])
    assert torch.equal(o1[0], o2[0])
    for a_, b_ in zip(g1, g2):
        assert relerr(a_, b_) < TOL[dtype] * 10
    o3, g3 = _grads(lambda: ref() * wgt.double(), [w, b, slope])
    assert relerr(o1[0], o3[0]) < TOL[dtype] * 5
    for a_, b_ in zip(g1, g3):
        assert relerr(a_, b_) < TOL[dtype] * 20
