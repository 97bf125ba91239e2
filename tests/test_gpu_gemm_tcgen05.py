"""The tcgen05/TMEM/TMA bf16 GEMM against an fp32 matmul of the same bf16-rounded operands."""
import pytest
import torch

from tests.emu_ops import EmuOps
from vrdone_b200.layout import PackLayout

pytestmark = pytest.mark.gpu

LENS = [37, 1, 2, 128, 5, 64, 93, 8, 3, 250, 300, 17]
TPADS = [128, 128, 128, 128, 128, 64, 96, 8, 128, 256, 512, 32]


@pytest.fixture(scope="module")
def ops():
    from vrdone_b200.cuda_ops import CudaOps
    return CudaOps()


def rnd(shape, seed, scale=1.0):
    return torch.randn(*shape, generator=torch.Generator().manual_seed(seed)) * scale


@pytest.mark.parametrize("taps,act,res,corr,N,K,odt", [
    (1, 0, 0, False, 512, 512, torch.float32),
    (1, 2, 0, False, 2048, 512, torch.bfloat16),
    (1, 0, 2, False, 512, 2048, torch.float32),
    (3, 0, 0, True, 512, 1024, torch.float32),
    (3, 0, 0, True, 512, 512, torch.float32),
    (1, 1, 1, False, 64, 256, torch.float32),
    (1, 0, 0, False, 192, 256, torch.float32),
    (1, 0, 1, False, 256, 256, torch.float32),
    (1, 0, 1, False, 512, 512, torch.float32),
    (1, 2, 0, False, 256, 256, torch.bfloat16),
    (1, 2, 0, False, 1024, 256, torch.bfloat16),
    (1, 0, 0, False, 256, 512, torch.float32),
    (1, 2, 0, False, 512, 1024, torch.bfloat16),     # 512-column tile (K >= 1024, bf16 out): GELU
    (1, 0, 0, False, 512, 1024, torch.bfloat16),
])
def test_gemm_bf16_tcgen05(ops, taps, act, res, corr, N, K, odt):
    lg, lc = PackLayout(LENS, TPADS, 4, "cuda"), PackLayout(LENS, TPADS, 4, "cpu")
    streams = 2
    M = streams * lc.levels[0].R
    a = rnd((M, K), 1).to(torch.bfloat16)
    a[(lc.levels[0].row_seq < 0).repeat(streams)] = 0
    w = rnd((N, taps * K), 2, K ** -0.5).to(torch.bfloat16)
    bias = rnd((N,), 3)
    r1 = rnd((M, N), 4) if res >= 1 else None
    r2 = rnd((M, N), 5) if res >= 2 else None
    cv = rnd((N,), 6) if corr else None
    ref = torch.empty(M, N)
    EmuOps().gemm(a, w, ref, bias=bias, taps=taps, act=act, res1=r1, res2=r2, corr=cv, lay=lc.levels[0], streams=streams)
    out = torch.full((M, N), 3.0, dtype=odt, device="cuda")
    cu = lambda t: None if t is None else t.cuda()
    ops.gemm(a.cuda(), w.cuda(), out, bias=bias.cuda(), taps=taps, act=act, res1=cu(r1), res2=cu(r2), corr=cu(cv), lay=lg.levels[0],
             streams=streams)
    torch.cuda.synchronize()
    err = float((out.float().cpu() - ref).abs().max())
    tol = 2e-3 if odt == torch.float32 else 2e-2      # fp32 accumulation order differs; bf16 output rounding
    assert err <= tol * float(ref.abs().max()), f"max abs err {err:.3e} vs scale {float(ref.abs().max()):.3e}"


def test_gemm_bf16_large_persistent(ops):
    """More tiles than SMs: exercises the persistent loop, the smem ring wrap-around and both TMEM stages."""
    M, N, K = 128 * 200, 512, 512
    a = rnd((M, K), 7).to(torch.bfloat16).cuda()
    w = rnd((N, K), 8, K ** -0.5).to(torch.bfloat16).cuda()
    out = torch.empty(M, N, device="cuda")
    ops.gemm(a, w, out)
    ref = a.float() @ w.float().t()
    torch.cuda.synchronize()
    assert float((out - ref).abs().max()) <= 2e-3 * float(ref.abs().max())


BIG_LENS = [int(x) for x in torch.randint(40, 400, (90,), generator=torch.Generator().manual_seed(5))]


@pytest.mark.parametrize("taps,act,res,corr,N,K,odt", [
    (1, 0, 2, False, 512, 512, torch.float32),
    (1, 2, 0, False, 2048, 512, torch.bfloat16),
    (1, 0, 1, False, 512, 2048, torch.float32),
    (3, 0, 0, True, 512, 1024, torch.float32),
    (1, 0, 0, False, 512, 512, torch.bfloat16),
    (1, 0, 1, False, 256, 512, torch.float32),
])
def test_gemm_bf16_cta_pairs(ops, taps, act, res, corr, N, K, odt):
    """M >= 16384 rows: the cta_group::2 path (a CTA pair per 256-row tile), odd number of 128-row blocks included."""
    tp = [512 if l > 256 else 256 for l in BIG_LENS]
    lg, lc = PackLayout(BIG_LENS, tp, 4, "cuda"), PackLayout(BIG_LENS, tp, 4, "cpu")
    streams = 1
    M = streams * lc.levels[0].R
    assert M >= 16384 and (M // 128) % 2 == 1, M
    a = rnd((M, K), 1).to(torch.bfloat16)
    a[(lc.levels[0].row_seq < 0).repeat(streams)] = 0
    w = rnd((N, taps * K), 2, K ** -0.5).to(torch.bfloat16)
    bias = rnd((N,), 3)
    r1 = rnd((M, N), 4) if res >= 1 else None
    r2 = rnd((M, N), 5) if res >= 2 else None
    cv = rnd((N,), 6) if corr else None
    ref = torch.empty(M, N)
    EmuOps().gemm(a, w, ref, bias=bias, taps=taps, act=act, res1=r1, res2=r2, corr=cv, lay=lc.levels[0], streams=streams)
    out = torch.full((M, N), 3.0, dtype=odt, device="cuda")
    cu = lambda t: None if t is None else t.cuda()
    ops.gemm(a.cuda(), w.cuda(), out, bias=bias.cuda(), taps=taps, act=act, res1=cu(r1), res2=cu(r2), corr=cu(cv), lay=lg.levels[0],
             streams=streams)
    torch.cuda.synchronize()
    err = float((out.float().cpu() - ref).abs().max())
    tol = 2e-3 if odt == torch.float32 else 2e-2
    assert err <= tol * float(ref.abs().max()), f"max abs err {err:.3e} vs scale {float(ref.abs().max()):.3e}"


@pytest.mark.parametrize("taps,K,relu,corr,big", [(3, 1024, True, True, False), (3, 512, True, True, True), (1, 512, False, False, False),
                                                  (1, 256, True, False, True)])
def test_gemm_layernorm_epilogue(ops, taps, K, relu, corr, big):
    """The embedding convs with their channel LayerNorm + ReLU as the GEMM's epilogue (N = 512, the tile spans the row), against the
    CPU emulation; ``big``: enough rows for the CTA-pair variant (M >= 16384)."""
    lens = LENS * 12 if big else LENS
    tpads = TPADS * 12 if big else TPADS
    lg, lc = PackLayout(lens, tpads, 4, "cuda"), PackLayout(lens, tpads, 4, "cpu")
    streams, N = 2, 512
    M = streams * lc.levels[0].R
    assert (M >= 16384) == big
    a = rnd((M, K), 1).to(torch.bfloat16)
    a[(lc.levels[0].row_seq < 0).repeat(streams)] = 0
    w = rnd((N, taps * K), 2, K ** -0.5).to(torch.bfloat16)
    bias, gm, be = rnd((N,), 3), rnd((N,), 7) * 0.2 + 1, rnd((N,), 8)
    cv = rnd((N,), 6) if corr else None
    ref = torch.empty(M, N, dtype=torch.bfloat16)
    EmuOps().gemm_ln(a, w, ref, (gm, be), bias=bias, taps=taps, corr=cv, relu=relu, lay=lc.levels[0], streams=streams)
    out = torch.full((M, N), 3.0, dtype=torch.bfloat16, device="cuda")
    ops.gemm_ln(a.cuda(), w.cuda(), out, (gm.cuda(), be.cuda()), bias=bias.cuda(), taps=taps, corr=None if cv is None else cv.cuda(),
                relu=relu, lay=lg.levels[0], streams=streams)
    torch.cuda.synchronize()
    ref = ref.float()
    err = float((out.float().cpu() - ref).abs().max())
    assert err <= 2e-2 * float(ref.abs().max()), f"max abs err {err:.3e} vs scale {float(ref.abs().max()):.3e}"
    # separator rows are written as zeros
    sep = (lc.levels[0].row_seq < 0).repeat(streams)
    assert float(out.float().cpu()[sep].abs().max()) == 0.0


@pytest.mark.parametrize("big", [False, True])
def test_gemm_residual_layernorm_epilogue(ops, big):
    """Attention output projection + residual (fp32) and the LayerNorm of the sum (bf16) from one call; ``big``: enough rows for the
    fused CTA-pair kernel (M >= 16384), otherwise the call runs the two launches itself."""
    lens = LENS * 12 if big else LENS
    tpads = TPADS * 12 if big else TPADS
    lg, lc = PackLayout(lens, tpads, 4, "cuda"), PackLayout(lens, tpads, 4, "cpu")
    streams, N, K = 2, 512, 512
    M = streams * lc.levels[0].R
    a = rnd((M, K), 1).to(torch.bfloat16)
    a[(lc.levels[0].row_seq < 0).repeat(streams)] = 0
    w = rnd((N, K), 2, K ** -0.5).to(torch.bfloat16)
    bias, gm, be = rnd((N,), 3), rnd((N,), 7) * 0.2 + 1, rnd((N,), 8)
    res = rnd((M, N), 4)
    ref, ref_ln = torch.empty(M, N), torch.empty(M, N, dtype=torch.bfloat16)
    EmuOps().gemm_res_ln(a, w, ref, ref_ln, (gm, be), bias=bias, res1=res, lay=lc.levels[0], streams=streams)
    out = torch.full((M, N), 3.0, device="cuda")
    out_ln = torch.full((M, N), 3.0, dtype=torch.bfloat16, device="cuda")
    ops.gemm_res_ln(a.cuda(), w.cuda(), out, out_ln, (gm.cuda(), be.cuda()), bias=bias.cuda(), res1=res.cuda(), lay=lg.levels[0],
                    streams=streams)
    torch.cuda.synchronize()
    err = float((out.cpu() - ref).abs().max())
    assert err <= 2e-3 * float(ref.abs().max()), f"sum: max abs err {err:.3e}"
    err = float((out_ln.float().cpu() - ref_ln.float()).abs().max())
    assert err <= 2e-2 * float(ref_ln.float().abs().max()), f"LayerNorm: max abs err {err:.3e}"
    sep = (lc.levels[0].row_seq < 0).repeat(streams)
    assert float(out.cpu()[sep].abs().max()) == 0.0 and float(out_ln.float().cpu()[sep].abs().max()) == 0.0
