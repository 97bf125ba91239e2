"""Per-kernel parity on the GPU: every C-ABI op against its CPU emulation (tests/emu_ops.py) on seeded ragged inputs."""
import pytest
import torch

from tests.emu_ops import EmuOps
from vrdone_b200.layout import PackLayout

pytestmark = pytest.mark.gpu

LENS = [37, 1, 2, 128, 5, 64, 93, 8, 3, 250]
TPADS = [128, 128, 128, 128, 128, 64, 96, 8, 128, 256]   # mixes "pad column exists" and "exactly full" at every level


@pytest.fixture(scope="module")
def ops():
    from vrdone_b200.cuda_ops import CudaOps
    return CudaOps()


@pytest.fixture(scope="module")
def lays():
    return PackLayout(LENS, TPADS, 4, "cuda"), PackLayout(LENS, TPADS, 4, "cpu")


def g(seed):
    return torch.Generator().manual_seed(seed)


def rnd(shape, seed, scale=1.0):
    return torch.randn(*shape, generator=g(seed)) * scale


def close(out_gpu, ref_cpu, tol, what=""):
    a, b = out_gpu.float().cpu(), ref_cpu.float()
    err = float((a - b).abs().max())
    ref = float(b.abs().max()) + 1e-12
    assert err <= tol * ref, f"{what}: max abs err {err:.3e} vs scale {ref:.3e} (tol {tol})"


def adt_tol(dt):
    return (torch.float32, 2e-5) if dt == "fp32" else (torch.bfloat16, 1.5e-2)


@pytest.mark.parametrize("token_major", [False, True])
def test_pack_pairs(ops, lays, token_major):
    lg, lc = lays
    nv, nc, nbs, nbe = 256, 128, 5, 8
    C = 2 * nv + 2 * nc + nbs + 2 * nbe
    feats = []
    for i, L in enumerate(LENS):
        if i % 2 == 0 or token_major:
            feats.append(rnd((L, C), 100 + i).permute(1, 0))          # (C, L) view of token-major memory (data loader)
        else:
            feats.append(rnd((C, L), 100 + i))                        # dense (C, L)
    R = lc.levels[0].R
    for dt in (torch.float32, torch.bfloat16):
        ref = [torch.empty(2 * R, nv, dtype=dt), torch.empty(2 * R, nc, dtype=dt), torch.empty(R, 8), torch.empty(2 * R, 8)]
        EmuOps().pack_pairs(feats, None, lc.levels[0], nv, nc, nbs, nbe, *ref)
        dev = [f.cuda() for f in feats]
        ptrs = torch.tensor([f.data_ptr() for f in dev], dtype=torch.int64).cuda()
        strides = torch.tensor([[f.stride(0), f.stride(1)] for f in dev], dtype=torch.int64).cuda()
        out = [torch.full_like(r, 7.0).cuda() for r in ref]
        ops.pack_pairs(ptrs, strides, lg.levels[0], nv, nc, nbs, nbe, *out, token_major=token_major)
        torch.cuda.synchronize()
        for o, r in zip(out, ref):
            assert torch.equal(o.cpu(), r)


@pytest.mark.parametrize("taps,act,res,corr,N,K", [(1, 0, 0, False, 512, 512), (1, 2, 0, False, 2048, 512), (1, 0, 2, False, 512, 2048),
                                                   (3, 0, 0, True, 512, 256), (1, 1, 1, False, 64, 256), (1, 0, 0, False, 144, 256)])
def test_gemm_fp32(ops, lays, taps, act, res, corr, N, K):
    lg, lc = lays
    streams = 2
    M = streams * lc.levels[0].R
    a = rnd((M, K), 1)
    a[(lc.levels[0].row_seq < 0).repeat(streams)] = 0
    w = rnd((N, taps * K), 2, K ** -0.5)
    bias = rnd((N,), 3)
    r1 = rnd((M, N), 4) if res >= 1 else None
    r2 = rnd((M, N), 5) if res >= 2 else None
    cv = rnd((N,), 6) if corr else None
    ref = torch.empty(M, N)
    EmuOps().gemm(a, w, ref, bias=bias, taps=taps, act=act, res1=r1, res2=r2, corr=cv, lay=lc.levels[0], streams=streams)
    out = torch.empty(M, N, device="cuda")
    cu = lambda t: None if t is None else t.cuda()
    ops.gemm(a.cuda(), w.cuda(), out, bias=bias.cuda(), taps=taps, act=act, res1=cu(r1), res2=cu(r2), corr=cu(cv), lay=lg.levels[0],
             streams=streams)
    # fp32 operands run on the tensor cores as 3 x bf16 split products (hi hi + lo hi + hi lo, fp32 accumulation): each operand keeps
    # 16 mantissa bits, so a product carries ~2^-16 relative error instead of fp32's 2^-24 (N = 144 does not fit the MMA tile and
    # takes the CUDA-core kernel)
    close(out, ref, 6e-5, "gemm fp32")


@pytest.mark.parametrize("dt", ["fp32", "bf16"])
def test_layernorm_and_small_conv(ops, lays, dt):
    lg, lc = lays
    odt, tol = adt_tol(dt)
    R = lc.levels[0].R
    x = rnd((2 * R, 512), 1, 2.0) + 0.5
    gm, be = rnd((512,), 2) * 0.3 + 1, rnd((512,), 3)
    ref = torch.empty(2 * R, 1024, dtype=odt)
    out = torch.zeros(2 * R, 1024, dtype=odt, device="cuda")
    EmuOps().layernorm(x, gm, be, ref[:, 512:], relu=True, lay=lc.levels[0], streams=2)
    ops.layernorm(x.cuda(), gm.cuda(), be.cuda(), out[:, 512:], relu=True, lay=lg.levels[0], streams=2)
    close(out[:, 512:], ref[:, 512:], tol, "layernorm")
    xb = rnd((2 * R, 8), 4)
    xb[(lc.levels[0].row_seq < 0).repeat(2)] = 0
    w, b = rnd((24, 512), 5, 0.2), rnd((512,), 6)
    EmuOps().small_conv(xb, 8, w, b, (gm, be), True, ref[:, :512], lc.levels[0], 2)
    ops.small_conv(xb.cuda(), 8, w.cuda(), b.cuda(), (gm.cuda(), be.cuda()), True, out[:, :512], lg.levels[0], 2)
    close(out[:, :512], ref[:, :512], tol, "small_conv+ln")
    w5 = rnd((15, 512), 7, 0.2)
    xb[:, 5:] = 0
    EmuOps().small_conv(xb[:R].contiguous(), 5, w5, b, None, False, ref[:R, :512], lc.levels[0], 1)
    ops.small_conv(xb[:R].contiguous().cuda(), 5, w5.cuda(), b.cuda(), None, False, out[:R, :512], lg.levels[0], 1)
    close(out[:R, :512], ref[:R, :512], tol, "small_conv")


@pytest.mark.parametrize("dt", ["fp32", "bf16"])
@pytest.mark.parametrize("stride,C,streams,flags", [(1, 512, 2, [True, True, True]), (1, 512, 1, [True, True, False]),
                                                    (2, 512, 1, [True, True, True]), (1, 256, 1, [False, False]),
                                                    (1, 512, 1, [True]), (1, 512, 1, [False, False])])
def test_dwconv_ln(ops, lays, dt, stride, C, streams, flags):
    lg, lc = lays
    odt, tol = adt_tol(dt)
    li, lo = (0, 0) if stride == 1 else (1, 2)
    x = rnd((streams * lc.levels[li].R, C), 1, 1.5)
    x[(lc.levels[li].row_seq < 0).repeat(streams)] = 0
    pre = (rnd((C,), 2) * 0.2 + 1, rnd((C,), 3))
    br_c, br_g = [], []
    outs_c, outs_g = [], []
    for b, use_pre in enumerate(flags):
        w, gm, be = rnd((3, C), 10 + b, 0.6), rnd((C,), 20 + b) * 0.2 + 1, rnd((C,), 30 + b)
        oc = torch.empty(streams * lc.levels[lo].R, C, dtype=odt)
        og = torch.full_like(oc, 3.0).cuda()
        outs_c.append(oc); outs_g.append(og)
        br_c.append((w, use_pre, gm, be, oc))
        br_g.append((w.cuda(), use_pre, gm.cuda(), be.cuda(), og))
    EmuOps().dwconv_ln(x, lc.levels[li], lc.levels[lo], stride, pre, br_c, streams)
    ops.dwconv_ln(x.cuda(), lg.levels[li], lg.levels[lo], stride, (pre[0].cuda(), pre[1].cuda()), br_g, streams)
    for og, oc in zip(outs_g, outs_c):
        close(og, oc, tol, "dwconv_ln")


@pytest.mark.parametrize("dt", ["fp32", "bf16"])
@pytest.mark.parametrize("n_head,w", [(8, 4), (4, 3)])
def test_window_and_full_attention(ops, lays, dt, n_head, w):
    lg, lc = lays
    adt, tol = adt_tol(dt)
    R = lc.levels[0].R
    q, k, v = [rnd((2 * R, 512), s, 0.5).to(adt) for s in (1, 2, 3)]
    ref = torch.empty(2 * R, 512, dtype=adt)
    out = torch.full((2 * R, 512), 5.0, dtype=adt, device="cuda")
    EmuOps().window_attn(q, k, v, ref, lc.levels[0], n_head, w, 2)
    ops.window_attn(q.cuda(), k.cuda(), v.cuda(), out, lg.levels[0], n_head, w, 2)
    valid2 = (lc.levels[0].row_seq >= 0).repeat(2)    # separator rows are don't-care for the only consumer (the projection GEMM)
    close(out[valid2], ref[valid2], tol, "window_attn")
    EmuOps().full_attn(q[:R], k[:R], v[:R], ref[:R], lc.levels[0], n_head)
    out.fill_(5.0)
    ops.full_attn(q[:R].cuda(), k[:R].cuda(), v[:R].cuda(), out[:R], lg.levels[0], n_head)
    valid = lc.levels[0].row_seq >= 0     # separator rows of the full-attention output are left unwritten (don't-care for proj)
    close(out[:R][valid], ref[:R][valid], tol, "full_attn")


def test_maxpool_and_fpn(ops, lays):
    lg, lc = lays
    e = [rnd((lc.levels[l].R, 512), 40 + l) for l in range(4)]
    for l in range(4):
        e[l][lc.levels[l].row_seq < 0] = 0
    ref = torch.empty(lc.levels[1].R, 512)
    out = torch.full_like(ref, 9.0).cuda()
    EmuOps().maxpool_skip(e[0], lc.levels[0], lc.levels[1], ref)
    ops.maxpool_skip(e[0].cuda(), lg.levels[0], lg.levels[1], out)
    assert torch.equal(out.cpu(), ref)
    pre = (rnd((512,), 1) * 0.2 + 1, rnd((512,), 2))
    ln = [(rnd((256,), 50 + l) * 0.2 + 1, rnd((256,), 60 + l)) for l in range(4)]
    wt = rnd((3, 512), 3, 0.5)
    y_c = torch.empty(lc.levels[3].R, 256)
    y_g = torch.full_like(y_c, 9.0).cuda()
    EmuOps().fpn_top(e[3], lc.levels[3], pre, wt, ln[3], y_c)
    ops.fpn_top(e[3].cuda(), lg.levels[3], (pre[0].cuda(), pre[1].cuda()), wt.cuda(), (ln[3][0].cuda(), ln[3][1].cuda()), y_g)
    close(y_g, y_c, 2e-5, "fpn_top")
    for l in (2, 1, 0):
        cur = rnd((lc.levels[l].R, 256), 70 + l)
        lat = (rnd((256,), 80 + l) * 0.2 + 1, rnd((256,), 90 + l))
        w = rnd((3, 256), 100 + l, 0.5)
        o_c = torch.empty(lc.levels[l].R, 256)
        o_g = torch.full_like(o_c, 9.0).cuda()
        EmuOps().fpn_level(cur, y_c, lc.levels[l], lc.levels[l + 1], lat, ln[l + 1][1], w, ln[l], o_c)
        ops.fpn_level(cur.cuda(), y_g, lg.levels[l], lg.levels[l + 1], (lat[0].cuda(), lat[1].cuda()), ln[l + 1][1].cuda(), w.cuda(),
                      (ln[l][0].cuda(), ln[l][1].cuda()), o_g)
        close(o_g, o_c, 5e-5, f"fpn_level{l}")
        y_c, y_g = o_c, o_g
    w, b = rnd((3, 256), 5, 0.5), rnd((256,), 6)
    m_c = torch.empty(lc.levels[0].R, 256)
    m_g = torch.full_like(m_c, 9.0).cuda()
    EmuOps().mask_features(y_c, lc.levels[0], ln[0][1], w, b, m_c)
    ops.mask_features(y_g, lg.levels[0], ln[0][1].cuda(), w.cuda(), b.cuda(), m_g)
    close(m_g, m_c, 5e-5, "mask_features")


@pytest.mark.parametrize("dt", ["fp32", "bf16"])
@pytest.mark.parametrize("Q,n_head", [(9, 8), (10, 4)])
def test_query_ops(ops, lays, dt, Q, n_head):
    lg, lc = lays
    adt, tol = adt_tol(dt)
    B = lc.B
    MQ = (B * Q + 127) // 128 * 128
    x = rnd((MQ, 256), 1)
    ln, ln2 = (rnd((256,), 2) * 0.2 + 1, rnd((256,), 3)), (rnd((256,), 4) * 0.2 + 1, rnd((256,), 5))
    pos, dw = rnd((Q, 256), 6), rnd((256,), 7)
    cu = lambda t: tuple(u.cuda() for u in t) if isinstance(t, tuple) else (None if t is None else t.cuda())
    for args in [(ln, pos, None, None), (None, None, None, None), (ln, pos, dw, ln2), (ln, None, None, None)]:
        ref = torch.empty(MQ, 256, dtype=adt)
        out = torch.full_like(ref, 4.0).cuda()
        EmuOps().query_ln(x, args[0], args[1], Q, B * Q, args[2], args[3], ref)
        ops.query_ln(x.cuda(), cu(args[0]), cu(args[1]), Q, B * Q, cu(args[2]), cu(args[3]), out)
        close(out, ref, tol, "query_ln")
    q, k, v = [rnd((MQ, 256), s, 0.4).to(adt) for s in (11, 12, 13)]
    ref = torch.empty(MQ, 256, dtype=adt)
    out = torch.zeros(MQ, 256, dtype=adt, device="cuda")
    EmuOps().query_self_attn(q, k, v, ref, B, Q, n_head)
    ops.query_self_attn(q.cuda(), k.cuda(), v.cuda(), out, B, Q, n_head)
    close(out, ref, tol, "query_self_attn")
    R3 = lc.levels[3].R
    kk, vv = [rnd((R3, 256), s, 0.4).to(adt) for s in (14, 15)]
    out.zero_()
    EmuOps().query_cross_attn(q, kk, vv, ref, lc.levels[3], Q, n_head)
    ops.query_cross_attn(q.cuda(), kk.cuda(), vv.cuda(), out, lg.levels[3], Q, n_head)
    close(out, ref, tol, "query_cross_attn")


@pytest.mark.parametrize("Q,n_cls,topk", [(9, 51, 6), (10, 133, 8), (9, 51, 1)])
def test_heads(ops, lays, Q, n_cls, topk):
    lg, lc = lays
    B = lc.B
    MQ = (B * Q + 127) // 128 * 128
    R0 = lc.levels[0].R
    me, mf = rnd((MQ, 256), 1, 0.3), rnd((R0, 256), 2, 0.3)
    masks_c, fl_c = torch.empty(R0, Q), torch.empty(B, Q, 2, dtype=torch.int32)
    masks_g, fl_g = torch.zeros(R0, Q, device="cuda"), torch.empty(B, Q, 2, dtype=torch.int32, device="cuda")
    EmuOps().mask_logits(me, mf, lc.levels[0], Q, masks_c, fl_c)
    ops.mask_logits(me.cuda(), mf.cuda(), lg.levels[0], Q, masks_g, fl_g)
    close(masks_g, masks_c, 2e-5, "mask_logits")
    # bit-exact binarisation given equal logits: recompute first/last from the GPU's own logits
    mg = masks_g.cpu()
    for i, L in enumerate(LENS):
        r0 = int(lc.levels[0].off[i])
        act = torch.sigmoid(mg[r0:r0 + L]) > 0.5
        for qi in range(Q):
            nz = torch.nonzero(act[:, qi]).flatten()
            exp = [int(nz[0]), int(nz[-1])] if nz.numel() else [-1, -1]
            assert fl_g[i, qi].tolist() == exp
    ncp = (n_cls + 15) // 16 * 16
    logits = rnd((MQ, ncp), 3, 2.0)
    sc_c, id_c = torch.empty(B * Q, topk), torch.empty(B * Q, topk, dtype=torch.int32)
    sc_g, id_g = torch.empty(B * Q, topk, device="cuda"), torch.empty(B * Q, topk, dtype=torch.int32, device="cuda")
    EmuOps().softmax_topk(logits, B * Q, n_cls, topk, sc_c, id_c)
    ops.softmax_topk(logits.cuda(), B * Q, n_cls, topk, sc_g, id_g)
    assert torch.equal(id_g.cpu(), id_c)
    close(sc_g, sc_c, 1e-5, "topk scores")
    ref_ids = torch.topk(torch.softmax(logits[:B * Q, :n_cls], -1)[:, 1:], topk, -1).indices + 1
    assert torch.equal(id_g.cpu().long(), ref_ids)


@pytest.mark.parametrize("scale", [0.5, 2.0])
def test_full_attention_tcgen05_long_and_ragged(ops, scale):
    """The tcgen05 / TMEM full-attention kernel (bf16, head_dim 64) on pairs from 1 frame to several query tiles and many key
    blocks (online-softmax rescaling of the TMEM accumulator, partially filled key blocks, tiles that start anywhere in the
    64-row scan blocks), against the fp32 emulation; larger score scale = peakier softmax = more rescaling."""
    lens = [1, 2, 63, 64, 65, 127, 128, 129, 200, 31, 513, 7, 700, 256, 3, 90]
    tp = [((l + 63) // 64) * 64 if l > 512 else 512 for l in lens]
    tp = [max(t, 64) for t in tp]
    lg, lc = PackLayout(lens, tp, 4, "cuda"), PackLayout(lens, tp, 4, "cpu")
    R = lc.levels[0].R
    q, k, v = [rnd((R, 512), s, scale if s < 3 else 1.0).to(torch.bfloat16) for s in (1, 2, 3)]
    ref = torch.empty(R, 512, dtype=torch.bfloat16)
    EmuOps().full_attn(q, k, v, ref, lc.levels[0], 8)
    out = torch.full((R, 512), 5.0, dtype=torch.bfloat16, device="cuda")
    ops.full_attn(q.cuda(), k.cuda(), v.cuda(), out, lg.levels[0], 8)
    torch.cuda.synchronize()
    valid = lc.levels[0].row_seq >= 0
    close(out[valid], ref[valid], 1.5e-2, "full_attn tcgen05")
    assert torch.all(out[~valid.cuda()] == 5.0), "separator rows must stay untouched"
    # a second launch on the same inputs is bit-identical (no dependence on scheduling)
    out2 = torch.full_like(out, 5.0)
    ops.full_attn(q.cuda(), k.cuda(), v.cuda(), out2, lg.levels[0], 8)
    assert torch.equal(out, out2)


def test_launcher_options_roundtrip(ops):
    """vrd_set_option / vrd_get_option: the experiment switches of the launchers; unknown names fail loudly."""
    for name in ("pdl", "dw_cfg", "gemm_spec", "embed_ln", "proj_ln"):
        cur = ops.get_option(name)
        assert ops.set_option(name, cur) == cur and ops.get_option(name) == cur
    with pytest.raises(ValueError):
        ops.set_option("no_such_option", 1)
    with pytest.raises(ValueError):
        ops.get_option("no_such_option")
