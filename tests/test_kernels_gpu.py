"""Per-kernel parity on the GPU: every C-ABI launcher against a plain torch fp32
restatement of the same arithmetic (the reference operators are cited in
include/gencast_b200.h)."""
import math
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _swish(x):
    return x * torch.sigmoid(x)


def _gelu_tanh(x):
    return 0.5 * x * (1 + torch.tanh(math.sqrt(2 / math.pi) * (x + 0.044715 * x ** 3)))


def _ln(x, eps=1e-6):
    mean = x.mean(-1, keepdim=True)
    var = ((x * x).mean(-1, keepdim=True) - mean * mean).clamp_min(0)
    return (x - mean) * torch.rsqrt(var + eps)


def _rel(a, b):
    return float((a.float() - b.float()).abs().max() / b.float().abs().max().clamp_min(1e-30))


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float32])
@pytest.mark.parametrize("m,n,ks", [(300, 128, (64,)), (1000, 512, (128, 192)), (4099, 256, (256, 64, 128)),
                                    (128, 2048, (512,)), (70000, 512, (512,))])
def test_gemm_plain(cuda_device, dtype, m, n, ks):
    from gencast_flax_nnx_b200 import ops
    g = torch.Generator(device="cpu").manual_seed(m + n)
    segs, ref = [], torch.zeros(m, n, dtype=torch.float64)
    for k in ks:
        a = torch.randn(m, k, generator=g).to(dtype)
        w = (torch.randn(n, k, generator=g) / math.sqrt(sum(ks))).to(dtype)
        ref += a.double() @ w.double().t()
        segs.append((a.to(cuda_device), w.to(cuda_device)))
    out = torch.empty(m, n, dtype=torch.float32, device=cuda_device)
    ops.gemm(segs, out)
    torch.cuda.synchronize()
    assert _rel(out.cpu(), ref) < (2e-5 if dtype == torch.float32 else 1e-4)


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float32])
@pytest.mark.parametrize("act", [None, "swish", "gelu_tanh"])
def test_gemm_fused_epilogue(cuda_device, dtype, act):
    from gencast_flax_nnx_b200 import ops
    m, n, k, ns, nr = 777, 256, 128, 200, 50
    g = torch.Generator(device="cpu").manual_seed(5)
    a = torch.randn(m, k, generator=g).to(dtype)
    w = (torch.randn(n, k, generator=g) / math.sqrt(k)).to(dtype)
    bias = torch.randn(n, generator=g)
    addend = torch.randn(m, n, generator=g).to(dtype)
    gs = torch.randn(ns, n, generator=g).to(dtype)
    gr = torch.randn(nr, n, generator=g).to(dtype)
    si = torch.randint(0, ns, (m,), generator=g, dtype=torch.int32)
    ri = torch.randint(0, nr, (m,), generator=g, dtype=torch.int32)
    res = torch.randn(m, n, generator=g)
    alpha = torch.tensor([0.37])
    pre = alpha.double() * (a.double() @ w.double().t()) + bias.double() + addend.double() + gs.double()[si.long()] + gr.double()[ri.long()]
    fn = {None: lambda x: x, "swish": _swish, "gelu_tanh": _gelu_tanh}[act]
    ref = fn(pre) + res.double()
    d = cuda_device
    for out_dtype, tol in ((torch.float32, 3e-5), (torch.bfloat16, 1e-2)):
        out = torch.empty(m, n, dtype=out_dtype, device=d)
        ops.gemm([(a.to(d), w.to(d))], out, bias=bias.to(d), act=act, addend=addend.to(d),
                 gathers=[(gs.to(d), si.to(d)), (gr.to(d), ri.to(d))], residual=res.to(d), alpha=alpha.to(d))
        torch.cuda.synchronize()
        assert _rel(out.cpu(), ref) < tol


@pytest.mark.parametrize("m,n,k", [(40000, 512, 256), (40000, 128, 128), (9000, 2048, 256), (5000, 1536, 128)])
def test_gemm_large_staged_store_paths(cuda_device, m, n, k):
    """Large problems take the persistent / CTA-pair kernels whose bias+activation epilogues are staged
    in shared memory and written with TMA stores (bf16 and fp32), and whose in-place fp32 residual
    update is a TMA reduce-add."""
    from gencast_flax_nnx_b200 import ops
    g = torch.Generator(device="cpu").manual_seed(n + k)
    a = torch.randn(m, k, generator=g).to(torch.bfloat16)
    w = (torch.randn(n, k, generator=g) / math.sqrt(k)).to(torch.bfloat16)
    bias = torch.randn(n, generator=g)
    d = cuda_device
    ref = a.double() @ w.double().t() + bias.double()
    out = torch.full((m, n), float("nan"), dtype=torch.bfloat16, device=d)
    ops.gemm([(a.to(d), w.to(d))], out, bias=bias.to(d), act="gelu_tanh")
    assert _rel(out.cpu(), _gelu_tanh(ref)) < 1e-2
    out32 = torch.full((m, n), float("nan"), dtype=torch.float32, device=d)
    ops.gemm([(a.to(d), w.to(d))], out32, bias=bias.to(d), act="swish")
    assert _rel(out32.cpu(), _swish(ref)) < 3e-3          # MUFU tanh in the bf16-path activations
    x0 = torch.randn(m, n, generator=g)
    x = x0.to(d).clone()
    ops.gemm([(a.to(d), w.to(d))], x, bias=bias.to(d), residual=x)      # x += a @ w^T + b, in place
    assert _rel(x.cpu(), ref + x0.double()) < 3e-5
    # view with a leading dimension larger than n
    big = torch.zeros(m, n + 64, dtype=torch.bfloat16, device=d)
    ops.gemm([(a.to(d), w.to(d))], big[:, :n], bias=bias.to(d))
    assert _rel(big[:, :n].cpu(), ref) < 1e-2 and torch.all(big[:, n:] == 0)


@pytest.mark.parametrize("m,n,k,two", [(40001, 512, 128, True), (20000, 256, 64, True), (30000, 512, 64, False),
                                       (40001, 384, 64, True)])
def test_gemm_large_staged_gathers(cuda_device, m, n, k, two):
    """Edge-MLP shape (common/typed_graph_net.py:134-159): e @ W1e + P_s[senders] + P_r[receivers], swish.
    At these sizes the CTA-pair / persistent kernels run and the row gathers go through the
    shared-memory transposition; m is not a multiple of the tile so the last tile is ragged."""
    from gencast_flax_nnx_b200 import ops
    g = torch.Generator(device="cpu").manual_seed(m + n)
    ns, nr = 5000, 1237
    a = torch.randn(m, k, generator=g).to(torch.bfloat16)
    w = (torch.randn(n, k, generator=g) / math.sqrt(k)).to(torch.bfloat16)
    bias = torch.randn(n, generator=g)
    gs = torch.randn(ns, n, generator=g).to(torch.bfloat16)
    gr = torch.randn(nr, n, generator=g).to(torch.bfloat16)
    si = torch.randint(0, ns, (m,), generator=g, dtype=torch.int32)
    ri = torch.sort(torch.randint(0, nr, (m,), generator=g, dtype=torch.int32)).values
    d = cuda_device
    pre = a.double() @ w.double().t() + bias.double() + gs.double()[si.long()]
    gathers = [(gs.to(d), si.to(d))]
    if two:
        pre = pre + gr.double()[ri.long()]
        gathers.append((gr.to(d), ri.to(d)))
    for out_dtype, tol in ((torch.bfloat16, 1e-2), (torch.float32, 3e-3)):
        out = torch.full((m, n), float("nan"), dtype=out_dtype, device=d)
        ops.gemm([(a.to(d), w.to(d))], out, bias=bias.to(d), act="swish", gathers=gathers)
        assert _rel(out.cpu(), _swish(pre)) < tol
    out = torch.full((m, n), float("nan"), dtype=torch.float32, device=d)
    ops.gemm([(a.to(d), w.to(d))], out, bias=bias.to(d), gathers=gathers)
    assert _rel(out.cpu(), pre) < 3e-5


@pytest.mark.parametrize("cols", [128, 256, 512])
@pytest.mark.parametrize("in_dtype,out_dtype", [(torch.float32, torch.float32), (torch.float32, torch.bfloat16),
                                                (torch.bfloat16, torch.float32)])
def test_ln_cond(cuda_device, cols, in_dtype, out_dtype):
    from gencast_flax_nnx_b200 import ops
    rows = 1001
    g = torch.Generator(device="cpu").manual_seed(cols)
    x = (torch.randn(rows, cols, generator=g) * 3 + 1).to(in_dtype)
    so = torch.cat([1 + 0.1 * torch.randn(cols, generator=g), torch.randn(cols, generator=g)])
    res = torch.randn(rows, cols, generator=g)
    d = cuda_device
    out = torch.empty(rows, cols, dtype=out_dtype, device=d)
    ops.ln_cond(x.to(d), out, so.to(d), residual=res.to(d))
    ref = _ln(x.double()) * so[:cols].double() + so[cols:].double() + res.double()
    assert _rel(out.cpu(), ref) < (1e-5 if out_dtype == torch.float32 else 1e-2)
    out2 = torch.empty(rows, cols, dtype=out_dtype, device=d)
    ops.ln_cond(x.to(d), out2, so.to(d), layer_norm=False)
    ref2 = x.double() * so[:cols].double() + so[cols:].double()
    assert _rel(out2.cpu(), ref2) < (1e-5 if out_dtype == torch.float32 else 1e-2)


@pytest.mark.parametrize("cols", [128, 256, 512])
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_segment_sum(cuda_device, cols, dtype):
    from gencast_flax_nnx_b200 import ops
    from gencast_flax_nnx_b200.graph import csr_by_receiver
    rng = np.random.default_rng(cols)
    nseg, nedge = 500, 6000
    recv = rng.integers(0, nseg, size=nedge)
    recv[:900] = 7            # one heavy receiver (block-cooperative path)
    recv[900:1100] = 123      # another in a different block
    recv = recv[recv != 11]   # one empty segment
    nedge = len(recv)
    row_ptr, perm = csr_by_receiver(recv, nseg)
    g = torch.Generator(device="cpu").manual_seed(1)
    y = torch.randn(nedge, cols, generator=g).to(dtype)
    so = torch.cat([1 + 0.1 * torch.randn(cols, generator=g), torch.randn(cols, generator=g)])
    d = cuda_device
    out = torch.full((nseg, cols), float("nan"), dtype=torch.float32, device=d)
    args = (y.to(d), out, so.to(d), torch.from_numpy(row_ptr).to(d), torch.from_numpy(perm).to(d))
    ops.ln_cond_segment_sum(*args)
    e = _ln(y.double()) * so[:cols].double() + so[cols:].double()
    ref = torch.zeros(nseg, cols, dtype=torch.float64).index_add_(0, torch.from_numpy(recv).long(), e)
    assert _rel(out.cpu(), ref) < 1e-5
    assert torch.all(out[11] == 0)
    # bitwise determinism
    out_b = torch.empty_like(out)
    ops.ln_cond_segment_sum(args[0], out_b, *args[2:])
    assert torch.equal(out, out_b)
    # already-sorted edges, no permutation, no LayerNorm (degree-3 mesh2grid shape)
    y3 = torch.randn(nseg * 3, cols, generator=g).to(dtype)
    rp3 = torch.arange(0, 3 * nseg + 1, 3, dtype=torch.int32)
    out3 = torch.empty(nseg, cols, dtype=torch.float32, device=d)
    ops.ln_cond_segment_sum(y3.to(d), out3, None, rp3.to(d), None, layer_norm=False)
    assert _rel(out3.cpu(), y3.double().reshape(nseg, 3, cols).sum(1)) < 1e-6


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("heads,head_dim", [(4, 32), (4, 64), (4, 128)])
def test_khop_attention(cuda_device, dtype, heads, head_dim):
    from gencast_flax_nnx_b200 import ops
    rng = np.random.default_rng(head_dim)
    n = 300
    hd = heads * head_dim
    g = torch.Generator(device="cpu").manual_seed(3)
    qkv = torch.randn(n, 3 * hd, generator=g).to(dtype)
    mask = rng.random((n, n)) < 0.2
    mask[np.arange(n), np.arange(n)] = True
    mask[5, :] = True            # a full row (> 32 * several chunks)
    ptr = np.zeros(n + 1, np.int32); ptr[1:] = np.cumsum(mask.sum(1))
    idx = np.nonzero(mask)[1].astype(np.int32)
    d = cuda_device
    out = torch.empty(n, hd, dtype=dtype, device=d)
    ops.khop_attention(qkv.to(d), out, torch.from_numpy(ptr).to(d), torch.from_numpy(idx).to(d), heads, head_dim)
    q, k, v = [t.double().reshape(n, heads, head_dim) for t in qkv.split(hd, dim=1)]
    logits = torch.einsum("qhd,khd->hqk", q, k) / math.sqrt(head_dim)
    logits = logits.masked_fill(~torch.from_numpy(mask)[None], float("-inf"))
    ref = torch.einsum("hqk,khd->qhd", torch.softmax(logits, -1), v).reshape(n, hd)
    assert _rel(out.cpu(), ref) < (1e-5 if dtype == torch.float32 else 1e-2)


@pytest.mark.parametrize("heads,head_dim", [(4, 64), (2, 128), (4, 128)])
@pytest.mark.parametrize("n,density,qk_scale", [(300, 0.2, 1.5), (1000, 0.05, 1.5), (128, 1.0, 1.5), (1000, 0.3, 5.0)])
def test_khop_attention_tensor_core(cuda_device, heads, head_dim, n, density, qk_scale):
    """Block-sparse tcgen05 attention against dense masked softmax attention in fp64."""
    from scipy import sparse
    from gencast_flax_nnx_b200 import ops
    from gencast_flax_nnx_b200.graph import khop_tiles
    rng = np.random.default_rng(head_dim + n)
    hd = heads * head_dim
    g = torch.Generator(device="cpu").manual_seed(4)
    # qk_scale 5 gives logits of order +-50: row maxima jump between key tiles, which exercises the
    # online-softmax offset update and the rescaling of O in tensor memory
    qkv = torch.randn(n, 3 * hd, generator=g)
    qkv[:, :2 * hd] *= qk_scale
    qkv[:, 2 * hd:] *= 1.5
    qkv = qkv.to(torch.bfloat16)
    mask = rng.random((n, n)) < density
    if n == 1000 and density < 0.1:
        mask[:, 300:700] = False         # empty key tiles for some query tiles (tile skipping)
        mask[400:, :200] = False
    if n == 1000 and density >= 0.1:
        # key sub-blocks of 32 that nobody attends to at the ends of key tiles (live key ranges of 1..3 sub-blocks)
        mask[:, 128:160] = False
        mask[:, 352:384] = False
        mask[:500, 256:320] = False
        mask[:, 640:736] = False
    mask[np.arange(n), np.arange(n)] = True
    tp, tk, tm = khop_tiles(sparse.csr_matrix(mask))
    d = cuda_device
    q, k, v = [t.double().reshape(n, heads, head_dim) for t in qkv.split(hd, dim=1)]
    logits = torch.einsum("qhd,khd->hqk", q, k) / math.sqrt(head_dim)
    logits = logits.masked_fill(~torch.from_numpy(mask)[None], float("-inf"))
    ref = torch.einsum("hqk,khd->qhd", torch.softmax(logits, -1), v).reshape(n, hd)
    from gencast_flax_nnx_b200.graph import pack_key_ranges
    outs = []
    for kv in (tk, pack_key_ranges(tk, tm)):          # plain list, and with the live key ranges annotated
        out = torch.full((n, hd), float("nan"), dtype=torch.bfloat16, device=d)
        ops.khop_attention_tiles(qkv.to(d), out, torch.from_numpy(tp).to(d), torch.from_numpy(kv).to(d),
                                 torch.from_numpy(tm.view(np.int32)).to(d), heads, head_dim)
        torch.cuda.synchronize()
        assert _rel(out.cpu(), ref) < 1.5e-2
        outs.append(out.cpu())
    assert _rel(outs[1], outs[0].double()) < 1e-2


@pytest.mark.parametrize("heads,head_dim", [(4, 64), (2, 128), (4, 128)])
@pytest.mark.parametrize("n,density,qk_scale", [(300, 0.2, 1.5), (1000, 0.05, 1.5), (128, 1.0, 1.5), (1000, 0.3, 5.0),
                                                (2300, 0.4, 1.5), (700, 0.1, 12.0)])
def test_khop_attention_gather(cuda_device, heads, head_dim, n, density, qk_scale):
    """tcgen05 attention over per-query-tile compacted key lists (gathered K / V tiles, single-pass online softmax)
    against dense masked softmax attention in fp64; bitwise repeatable."""
    from scipy import sparse
    from gencast_flax_nnx_b200 import ops
    from gencast_flax_nnx_b200.graph import khop_compact_steps
    rng = np.random.default_rng(head_dim + n)
    hd = heads * head_dim
    g = torch.Generator(device="cpu").manual_seed(4)
    # qk_scale 5 / 12 give logits of order +-50 / +-300 (x log2 e / sqrt d): row maxima jump between steps by far more
    # than 2^40, which exercises the overflow guard, the offset raise and the rescaling of O in tensor memory
    qkv = torch.randn(n, 3 * hd, generator=g)
    qkv[:, :2 * hd] *= qk_scale
    qkv[:, 2 * hd:] *= 1.5
    qkv = qkv.to(torch.bfloat16)
    mask = rng.random((n, n)) < density
    if n == 1000 and density < 0.1:
        mask[:, 300:700] = False         # keys nobody attends to: absent from every compacted list
        mask[400:, :200] = False
    if n == 700:
        mask[:, :] = np.triu(mask, 0)    # rows meet their first neighbour at different steps
        mask[100:200, :] = False         # rows with the diagonal only
    if n == 2300:
        mask[:128, :] = True             # one query tile attends to everything: 36 steps (> the staged key list)
    mask[np.arange(n), np.arange(n)] = True
    sp, keys, cm, work = khop_compact_steps(sparse.csr_matrix(mask))
    d = cuda_device
    q, k, v = [t.double().reshape(n, heads, head_dim) for t in qkv.split(hd, dim=1)]
    logits = torch.einsum("qhd,khd->hqk", q, k) / math.sqrt(head_dim)
    logits = logits.masked_fill(~torch.from_numpy(mask)[None], float("-inf"))
    ref = torch.einsum("hqk,khd->qhd", torch.softmax(logits, -1), v).reshape(n, hd)
    args = [torch.from_numpy(a.view(np.int32).reshape(-1) if a.dtype == np.uint32 else a).to(d) for a in (sp, keys, cm, work)]
    outs = []
    for _ in range(2):
        out = torch.full((n, hd), float("nan"), dtype=torch.bfloat16, device=d)
        ops.khop_attention_gather(qkv.to(d), out, *args, heads, head_dim)
        torch.cuda.synchronize()
        assert _rel(out.cpu(), ref) < 1.5e-2
        outs.append(out.cpu())
    assert torch.equal(outs[0], outs[1])


def test_khop_attention_gather_members_share_masks(cuda_device):
    """Two members evaluated together: key lists offset per member, one copy of the masks (mask_period)."""
    from scipy import sparse
    from gencast_flax_nnx_b200 import ops
    from gencast_flax_nnx_b200.graph import khop_compact_steps
    rng = np.random.default_rng(5)
    n, heads, head_dim = 256, 4, 64
    hd = heads * head_dim
    mask = rng.random((n, n)) < 0.15
    mask[np.arange(n), np.arange(n)] = True
    sp, keys, cm, work = khop_compact_steps(sparse.csr_matrix(mask))
    ns, nq = int(sp[-1]), len(sp) - 1
    d = cuda_device
    qkv = (torch.randn(2 * n, 3 * hd, generator=torch.Generator().manual_seed(1)) * 1.2).to(torch.bfloat16).to(d)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(d)
    single = []
    for b in range(2):
        out = torch.empty(n, hd, dtype=torch.bfloat16, device=d)
        ops.khop_attention_gather(qkv[b * n:(b + 1) * n], out, t(sp), t(keys), t(cm.view(np.int32).reshape(-1)), t(work), heads, head_dim)
        single.append(out)
    sp2 = np.concatenate([sp[:-1], sp[:-1] + ns, [2 * ns]]).astype(np.int32)
    keys2 = np.concatenate([keys, keys + n]).astype(np.int32)
    work2 = np.argsort(-np.diff(sp2), kind="stable").astype(np.int32)
    out2 = torch.empty(2 * n, hd, dtype=torch.bfloat16, device=d)
    ops.khop_attention_gather(qkv, out2, t(sp2), t(keys2), t(cm.view(np.int32).reshape(-1)), t(work2), heads, head_dim, mask_period=ns)
    torch.cuda.synchronize()
    assert torch.equal(out2[:n], single[0]) and torch.equal(out2[n:], single[1])


@pytest.mark.parametrize("cols,nrecv,members", [(128, 37, 1), (128, 50, 3), (256, 1000, 2), (512, 6000, 1), (512, 20000, 4),
                                                (512, 10512, 2)])
@pytest.mark.parametrize("out_dtype", [torch.bfloat16, torch.float32])
@pytest.mark.parametrize("pair", [False, True])
def test_edge_mlp_sum3_fused(cuda_device, cols, nrecv, members, out_dtype, pair, monkeypatch):
    """Fused degree-3 edge update + aggregation (gather-add-swish -> tcgen05 GEMM -> LayerNorm -> 3-row segment sum ->
    conditional affine) against the same chain in fp64 on the bf16-rounded operands; bitwise repeatable.
    pair: the opt-in column-split CTA-pair variant (GENCAST_EDGE_PAIR=1; used when there is no receiver table, cols >= 256
    and the period allows TMA-fed operands)."""
    from gencast_flax_nnx_b200 import ops
    if pair and cols < 256:
        pytest.skip("the pair variant needs cols >= 256")
    monkeypatch.setenv("GENCAST_EDGE_PAIR", "1" if pair else "0")
    g = torch.Generator(device="cpu").manual_seed(cols + nrecv)
    d = cuda_device
    R = nrecv * members
    E = 3 * R
    period = 3 * nrecv                                        # base rows are shared by the members
    n_s, n_r = 777, R
    bf = lambda t: t.to(torch.bfloat16)
    base = bf(torch.randn(period, cols, generator=g))
    gs = bf(torch.randn(n_s, cols, generator=g))
    gr = bf(torch.randn(n_r, cols, generator=g))
    idx_s = torch.randint(0, n_s, (E,), generator=g, dtype=torch.int32)
    idx_r = torch.arange(R, dtype=torch.int32).repeat_interleave(3)
    w2 = bf(torch.randn(cols, cols, generator=g) / math.sqrt(cols))
    b2 = torch.randn(cols, generator=g) * 0.1
    so = torch.cat([1 + 0.1 * torch.randn(cols, generator=g), torch.randn(cols, generator=g)])
    e = torch.arange(E)
    h = base.double()[e % period] + gs.double()[idx_s.long()] + gr.double()[idx_r.long()]
    h = bf((h * torch.sigmoid(h)).float()).double()
    y = h @ w2.double().T + b2.double()
    mean = y.mean(-1, keepdim=True)
    var = ((y * y).mean(-1, keepdim=True) - mean * mean).clamp_min(0)
    ln = (y - mean) / torch.sqrt(var + 1e-6)
    ref = ln.reshape(R, 3, cols).sum(1) * so[:cols].double() + 3 * so[cols:].double()
    outs = []
    for _ in range(2):
        out = torch.full((R, cols), float("nan"), dtype=out_dtype, device=d)
        ops.edge_mlp_sum3(base.to(d), [(gs.to(d), idx_s.to(d)), (gr.to(d), idx_r.to(d))], w2.to(d), b2.to(d), so.to(d), out)
        torch.cuda.synchronize()
        outs.append(out.cpu())
    # the hidden layer is rounded to bf16 on both sides; MUFU swish differs from the exact one by < 1 bf16 ulp
    assert _rel(outs[0], ref) < (1.2e-2 if out_dtype == torch.bfloat16 else 8e-3)
    assert torch.equal(outs[0], outs[1])
    # no receiver table = "receiver v's own row" (base / receiver rows by TMA when the period allows): the same bits
    out = torch.full((R, cols), float("nan"), dtype=out_dtype, device=d)
    ops.edge_mlp_sum3(base.to(d), [(gs.to(d), idx_s.to(d)), (gr.to(d), None)], w2.to(d), b2.to(d), so.to(d), out)
    torch.cuda.synchronize()
    if pair:
        # column-split CTA pair: the row statistics are summed in a different order
        assert _rel(out.cpu(), outs[0].double()) < 4e-3
    else:
        assert torch.equal(out.cpu(), outs[0])
    # without LayerNorm / affine / bias: plain sum of the three second-layer outputs
    out = torch.empty(R, cols, dtype=torch.float32, device=d)
    ops.edge_mlp_sum3(base.to(d), [(gs.to(d), idx_s.to(d)), (gr.to(d), idx_r.to(d))], w2.to(d), None, None, out, layer_norm=False)
    assert _rel(out.cpu(), (h @ w2.double().T).reshape(R, 3, cols).sum(1)) < 8e-3


@pytest.mark.parametrize("cols,period,members", [(128, 100, 1), (256, 1000, 3), (512, 5003, 2), (512, 128 * 7, 4)])
def test_edge_mlp_rows_fused(cuda_device, cols, period, members):
    """Tabulated first layer + gather + swish -> second layer (tcgen05) -> bf16 rows in one kernel against the same chain in
    fp64 on the bf16-rounded operands; member blocks that are not a whole number of 128-edge tiles; bitwise repeatable."""
    from gencast_flax_nnx_b200 import ops
    g = torch.Generator(device="cpu").manual_seed(cols + period)
    d = cuda_device
    E = period * members
    n_s = 611
    bf = lambda t: t.to(torch.bfloat16)
    base = bf(torch.randn(period, cols, generator=g))
    gs = bf(torch.randn(n_s, cols, generator=g))
    idx_s = torch.randint(0, n_s, (E,), generator=g, dtype=torch.int32)
    w2 = bf(torch.randn(cols, cols, generator=g) / math.sqrt(cols))
    b2 = torch.randn(cols, generator=g) * 0.1
    h = base.double()[torch.arange(E) % period] + gs.double()[idx_s.long()]
    h = bf((h * torch.sigmoid(h)).float()).double()
    ref = h @ w2.double().T + b2.double()
    outs = []
    for _ in range(2):
        out = torch.full((E, cols), float("nan"), dtype=torch.bfloat16, device=d)
        ops.edge_mlp_rows(base.to(d), (gs.to(d), idx_s.to(d)), w2.to(d), b2.to(d), out)
        torch.cuda.synchronize()
        outs.append(out.cpu())
    assert _rel(outs[0], ref) < 1.2e-2
    assert torch.equal(outs[0], outs[1])
    # against the two-kernel path it replaces (gc_edge_hidden -> gc_gemm): same hidden layer, same products
    e_h = torch.empty(E, cols, dtype=torch.bfloat16, device=d)
    e_y = torch.empty(E, cols, dtype=torch.bfloat16, device=d)
    ops.edge_hidden(base.to(d), [(gs.to(d), idx_s.to(d))], e_h, act="swish")
    ops.gemm([(e_h, w2.to(d))], e_y, bias=b2.to(d))
    assert _rel(outs[0], e_y.cpu().double()) < 4e-3
    # per-row LayerNorm statistics from the fp32 accumulator: {sum, sum of squares} of each half of the columns ...
    stats = torch.full((E, 4), float("nan"), dtype=torch.float32, device=d)
    out_s = torch.empty(E, cols, dtype=torch.bfloat16, device=d)
    ops.edge_mlp_rows(base.to(d), (gs.to(d), idx_s.to(d)), w2.to(d), b2.to(d), out_s, row_stats=stats)
    torch.cuda.synchronize()
    assert torch.equal(out_s.cpu(), outs[0])
    h2 = cols // 2
    want = torch.stack([ref[:, :h2].sum(1), (ref[:, :h2] ** 2).sum(1), ref[:, h2:].sum(1), (ref[:, h2:] ** 2).sum(1)], dim=1)
    got = stats.cpu().double()
    assert torch.isfinite(got).all()
    assert ((got - want).abs() / (want.abs() + 1.0 * math.sqrt(cols))).max() < 2e-3
    # ... which the segment sum consumes instead of reducing the rounded rows itself: same LayerNorm up to bf16 noise
    nseg = 97
    cuts = torch.sort(torch.randint(0, E + 1, (nseg - 1,), generator=g)).values
    rp = torch.cat([torch.zeros(1, dtype=torch.int64), cuts, torch.tensor([E])]).to(torch.int32).to(d)
    so = torch.cat([1 + 0.1 * torch.randn(cols, generator=g), torch.randn(cols, generator=g)]).to(d)
    agg_a = torch.empty(nseg, cols, dtype=torch.float32, device=d)
    agg_b = torch.empty(nseg, cols, dtype=torch.float32, device=d)
    ops.ln_cond_segment_sum(out_s, agg_a, so, rp, None, irregular=True)
    ops.ln_cond_segment_sum(out_s, agg_b, so, rp, None, irregular=True, row_stats=stats)
    torch.cuda.synchronize()
    assert _rel(agg_b.cpu(), agg_a.cpu().double()) < 5e-3


@pytest.mark.parametrize("cols,rows", [(128, 100), (256, 5000), (512, 20001), (512, 128)])
@pytest.mark.parametrize("res_dtype,out_dtype", [(None, torch.bfloat16), (torch.bfloat16, torch.bfloat16), (torch.bfloat16, torch.float32),
                                                 (torch.float32, torch.float32)])
def test_linear_ln_cond_fused(cuda_device, cols, rows, res_dtype, out_dtype):
    """Second MLP layer + LayerNorm + conditional affine + residual in one kernel against fp64; bitwise repeatable."""
    from gencast_flax_nnx_b200 import ops
    g = torch.Generator(device="cpu").manual_seed(cols + rows)
    d = cuda_device
    a = torch.randn(rows, cols, generator=g).to(torch.bfloat16)
    w = (torch.randn(cols, cols, generator=g) / math.sqrt(cols)).to(torch.bfloat16)
    b = torch.randn(cols, generator=g) * 0.1
    so = torch.cat([1 + 0.1 * torch.randn(cols, generator=g), torch.randn(cols, generator=g)])
    res = None if res_dtype is None else torch.randn(rows, cols, generator=g).to(res_dtype)
    y = a.double() @ w.double().T + b.double()
    mean = y.mean(-1, keepdim=True)
    var = ((y * y).mean(-1, keepdim=True) - mean * mean).clamp_min(0)
    ref = (y - mean) / torch.sqrt(var + 1e-6) * so[:cols].double() + so[cols:].double()
    if res is not None:
        ref = ref + res.double()
    outs = []
    for _ in range(2):
        out = torch.full((rows, cols), float("nan"), dtype=out_dtype, device=d)
        ops.linear_ln_cond(a.to(d), w.to(d), b.to(d), so.to(d), out, residual=None if res is None else res.to(d))
        torch.cuda.synchronize()
        outs.append(out.cpu())
    assert _rel(outs[0], ref) < (6e-3 if out_dtype == torch.bfloat16 else 2e-5 * 50)
    assert torch.equal(outs[0], outs[1])


def test_cond_tables_and_fold(cuda_device):
    from gencast_flax_nnx_b200 import ops
    g = torch.Generator(device="cpu").manual_seed(9)
    nfreq, layers, width = 32, 5, 256
    sigma = torch.tensor([80.0, 1.0, 0.03, 1e-6, 7.5])
    w0 = torch.randn(2 * nfreq, 32, generator=g) / 8; b0 = torch.randn(32, generator=g) * 0.1
    w1 = torch.randn(32, 16, generator=g) / 5.6; b1 = torch.randn(16, generator=g) * 0.1
    wc = torch.randn(layers, 16, 2 * width, generator=g) * 0.1; bc = torch.randn(layers, 2 * width, generator=g) * 0.1
    d = cuda_device
    table = torch.empty(len(sigma), layers, 2 * width, device=d)
    ops.cond_tables(sigma.to(d), w0.to(d), b0.to(d), w1.to(d), b1.to(d), 16.0, nfreq, wc.to(d), bc.to(d), table)
    z = torch.log(sigma.double())
    ang = z[:, None] * (2 * math.pi * torch.arange(1, nfreq + 1).double() / 16.0)
    f = torch.cat([torch.cos(ang), torch.sin(ang)], -1)
    cond = _gelu_tanh(f @ w0.double() + b0.double()) @ w1.double() + b1.double()
    ref = torch.einsum("sc,lcw->slw", cond, wc.double()) + bc.double()
    ref[..., :width] += 1
    assert _rel(table.cpu(), ref) < 2e-4       # fp32 sin/cos of |angle| up to ~170 rad

    for dtype in (torch.float32, torch.bfloat16):
        n, k = 256, 256
        w = (torch.randn(n, k, generator=g) / 16).to(dtype)
        bias = torch.randn(n, generator=g)
        so = torch.cat([1 + 0.1 * torch.randn(k, generator=g), torch.randn(k, generator=g)])
        w_out = torch.empty(n, k, dtype=dtype, device=d); b_out = torch.empty(n, device=d)
        ops.fold_affine_into_linear(w.to(d), bias.to(d), so.to(d), w_out, b_out)
        assert _rel(w_out.cpu(), w.double() * so[:k].double()) < (1e-6 if dtype == torch.float32 else 1e-2)
        assert _rel(b_out.cpu(), bias.double() + w.double() @ so[k:].double()) < 1e-5


def test_cond_tables_match_reference_fourier_features_mlp(cuda_device):
    """gc_cond_tables against the output of the reference's own FourierFeaturesMLP code (common/mlp.py:255-265,
    common/model_utils.py:728-757; fixture tests/golden/refshim_sampler.npz, tools/make_sampler_golden.py): with an
    identity conditional linear the table rows are (1 + cond[:8] | cond[8:])."""
    import os
    from gencast_flax_nnx_b200 import ops
    gold = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "refshim_sampler.npz"))
    d = cuda_device
    f32 = lambda k: torch.from_numpy(gold[k].astype(np.float32)).to(d)
    sigma = f32("encoder/sigmas")
    wc = torch.eye(16, device=d).reshape(1, 16, 16).contiguous()
    bc = torch.zeros(1, 16, device=d)
    table = torch.empty(len(sigma), 1, 16, device=d)
    ops.cond_tables(sigma, f32("encoder/linear_0/kernel"), f32("encoder/linear_0/bias"), f32("encoder/linear_1/kernel"),
                    f32("encoder/linear_1/bias"), 16.0, 32, wc, bc, table)
    got = table[:, 0].cpu().double()
    got[:, :8] -= 1.0
    assert _rel(got, torch.from_numpy(gold["encoder/cond"])) < 2e-4      # fp32 sin / cos of |angle| up to ~170 rad


def test_dpm_update_cast_pad_accumulate(cuda_device):
    from gencast_flax_nnx_b200 import ops
    g = torch.Generator(device="cpu").manual_seed(2)
    rows, cols = 1000, 82
    f = torch.randn(rows, 128, generator=g); x = torch.randn(rows, cols, generator=g); xb = torch.randn(rows, cols, generator=g)
    sched = torch.tensor([0.7, 0.3, 0.25, 1.7])
    d = cuda_device
    for xin_dtype in (torch.float32, torch.bfloat16):
        x_out = torch.empty(rows, cols, device=d)
        xin = torch.zeros(rows, 128, dtype=xin_dtype, device=d)
        ops.dpm_update(f.to(d), x.to(d), xb.to(d), sched.to(d), x_out, xin, cols)
        den = 0.7 * f[:, :cols] + 0.3 * x
        ref = 0.25 * xb + 0.75 * den
        assert _rel(x_out.cpu(), ref) < 1e-6
        assert _rel(xin.cpu()[:, :cols], 1.7 * ref) < (1e-6 if xin_dtype == torch.float32 else 1e-2)
        assert torch.all(xin[:, cols:] == 0)
    dst = torch.full((rows, 128), 5.0, dtype=torch.bfloat16, device=d)
    ops.cast_pad(x.to(d), dst)
    assert _rel(dst.cpu()[:, :cols], x) < 1e-2 and torch.all(dst[:, cols:] == 0)
    tot = torch.zeros(rows, cols, device=d); tot2 = torch.zeros(rows, cols, device=d)
    for _ in range(3):
        ops.ensemble_accumulate(x.to(d), tot, tot2)
    assert _rel(tot.cpu(), 3 * x) < 1e-6 and _rel(tot2.cpu(), 3 * x * x) < 1e-6


def test_errors_are_reported(cuda_device):
    from gencast_flax_nnx_b200 import ops, _lib
    d = cuda_device
    a = torch.zeros(10, 100, dtype=torch.bfloat16, device=d)     # k not a multiple of 64
    w = torch.zeros(128, 100, dtype=torch.bfloat16, device=d)
    out = torch.zeros(10, 128, device=d)
    with pytest.raises(_lib.GencastKernelError):
        ops.gemm([(a[:, :96], w[:, :96])], out)
    with pytest.raises(ValueError):
        ops.gemm([(a.cpu(), w.cpu())], out)


@pytest.mark.parametrize("M", [2, 5, 32])
def test_fair_crps_kernel(cuda_device, M):
    """gc_fair_crps + gc_column_sums against the defining formula (parallel.py docstring) in fp64, through the
    same entry point the ensemble code uses (parallel.fair_crps on one rank)."""
    from gencast_flax_nnx_b200 import ops, parallel
    g = torch.Generator(device="cpu").manual_seed(M)
    G, C = 777, 5
    x = torch.randn(M, G, C, generator=g)
    x[:, 10] = x[0, 10]                                   # ties
    y = torch.randn(G, C, generator=g)
    w = torch.rand(G, generator=g) + 0.1
    xd, yd = x.double(), y.double()
    skill = (xd - yd[None]).abs().mean(0)
    pair = (xd[:, None] - xd[None, :]).abs().sum((0, 1)) / 2 / (M * (M - 1))
    ref_pt = skill - pair
    d = cuda_device
    got_pt = ops.fair_crps(x.reshape(M, -1).to(d), y.reshape(-1).to(d), None, C).cpu().reshape(G, C)
    assert float((got_pt.double() - ref_pt).abs().max()) < 2e-5
    ref = (ref_pt * w.double()[:, None]).sum(0) / w.double().sum()
    got = parallel.fair_crps(x.to(d), y.to(d), w.to(d)).cpu()
    assert float((got.double() - ref).abs().max() / ref.abs().max()) < 1e-5
    again = parallel.fair_crps(x.to(d), y.to(d), w.to(d)).cpu()
    assert torch.equal(got, again)                       # fixed summation order
    host = parallel.fair_crps(x, y, w)                   # the torch formula used without a GPU
    assert float((host.double() - ref).abs().max() / ref.abs().max()) < 1e-5


@pytest.mark.parametrize("cols", [128, 256, 512])
@pytest.mark.parametrize("act", ["swish", None])
def test_edge_hidden(cuda_device, cols, act):
    """gc_edge_hidden: act(base[e % period] + g0[idx0[e]] + g1[idx1[e]]) -- the first edge-MLP layer
    (common/typed_graph_net.py:134-159) once its edge-feature part is tabulated per noise level."""
    from gencast_flax_nnx_b200 import ops
    g = torch.Generator(device="cpu").manual_seed(cols)
    period, members, ns, nr = 1003, 3, 400, 77
    rows = period * members
    base = torch.randn(period, cols, generator=g).to(torch.bfloat16)
    g0 = torch.randn(ns, cols, generator=g).to(torch.bfloat16)
    g1 = torch.randn(nr, cols, generator=g).to(torch.bfloat16)
    i0 = torch.randint(0, ns, (rows,), generator=g, dtype=torch.int32)
    i1 = torch.randint(0, nr, (rows,), generator=g, dtype=torch.int32)
    d = cuda_device
    fn = _swish if act == "swish" else (lambda x: x)
    pre = base.double().repeat(members, 1) + g0.double()[i0.long()] + g1.double()[i1.long()]
    out = torch.full((rows, cols), float("nan"), dtype=torch.bfloat16, device=d)
    ops.edge_hidden(base.to(d), [(g0.to(d), i0.to(d)), (g1.to(d), i1.to(d))], out, act=act)
    assert _rel(out.cpu(), fn(pre)) < 1e-2
    out1 = torch.full((rows, cols), float("nan"), dtype=torch.bfloat16, device=d)
    ops.edge_hidden(base.to(d), [(g0.to(d), i0.to(d))], out1, act=act)
    assert _rel(out1.cpu(), fn(base.double().repeat(members, 1) + g0.double()[i0.long()])) < 1e-2
