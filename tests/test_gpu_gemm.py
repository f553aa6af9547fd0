"""tcgen05 GEMM engine vs (a) the in-library SIMT cross-check kernel and (b) torch fp32 matmul on
the same bf16/tf32-rounded operands.  Covers all operand majors, ragged M/N/K, split-K, epilogue
options."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _mk(rows, cols, dtype, major, gen, pad=0):
    """logical [rows, K=cols] operand stored K-major ([rows, cols]) or MN-major ([cols, rows])."""
    shape = (rows, cols) if major == 0 else (cols, rows)
    full = torch.randn(shape[0], shape[1] + pad, generator=gen, device="cuda", dtype=torch.float32)
    full = full.to(dtype)
    view = full[:, : shape[1]]
    logical = view.float() if major == 0 else view.float().t()
    return view, logical


CASES = [
    # kind, a_major, b_major, M, N, K, tile_n, split_k
    (torch.bfloat16, 0, 0, 128, 128, 64, 128, 1),
    (torch.bfloat16, 0, 0, 128, 256, 512, 256, 1),
    (torch.bfloat16, 0, 0, 200, 9488, 512, 0, 1),
    (torch.bfloat16, 0, 0, 1024, 3072, 1024, 0, 1),
    (torch.bfloat16, 0, 0, 50, 2560, 1000, 64, 1),
    (torch.bfloat16, 0, 1, 300, 512, 9488, 0, 1),
    (torch.bfloat16, 1, 1, 2560, 512, 1111, 0, 1),
    (torch.bfloat16, 1, 1, 512, 2048, 4000, 128, 4),
    (torch.bfloat16, 1, 0, 256, 192, 320, 64, 1),
    (torch.bfloat16, 0, 0, 1024, 9488, 512, 192, 1),
    (torch.bfloat16, 0, 1, 1000, 1024, 3072, 192, 1),
    (torch.bfloat16, 1, 1, 9488, 512, 2000, 192, 2),
    (torch.float32, 0, 0, 130, 1024, 2048, 0, 1),
    (torch.float32, 0, 0, 96, 200, 300, 128, 1),
    (torch.float32, 0, 0, 512, 2048, 776, 0, 2),
]


@pytest.mark.parametrize("dtype,am,bm,M,N,K,tile_n,split_k", CASES)
def test_gemm_matches_simt_and_torch(dtype, am, bm, M, N, K, tile_n, split_k):
    from cooperativeimagecaptioning_b200 import ops
    gen = torch.Generator(device="cuda").manual_seed(1000 + M + N + K)
    pad = 8  # exercise lda != cols
    A, Al = _mk(M, K, dtype, am, gen, pad)
    B, Bl = _mk(N, K, dtype, bm, gen, pad)
    bias = torch.randn(N, generator=gen, device="cuda")
    ref = Al @ Bl.t() * 0.5 + bias

    mode = 2 if split_k > 1 else 0
    out = torch.zeros(M, N, device="cuda") if split_k > 1 else torch.full((M, N), float("nan"), device="cuda")
    if split_k > 1:
        # bias would be added once per split; fold it in afterwards
        ops.gemm(A, B, M, N, K, a_major=am, b_major=bm, alpha=0.5, mode=mode, out=out,
                 split_k=split_k, tile_n=tile_n)
        out += bias
    else:
        ops.gemm(A, B, M, N, K, a_major=am, b_major=bm, alpha=0.5, bias=bias, out=out,
                 tile_n=tile_n)
    simt = torch.empty(M, N, device="cuda")
    ops.gemm(A, B, M, N, K, a_major=am, b_major=bm, alpha=0.5, bias=bias, out=simt, backend=1)
    torch.cuda.synchronize()
    scale = ref.abs().max().item()
    # SIMT kernel is fp32 on the same operands -> agrees with torch to fp32 rounding
    assert (simt - ref).abs().max().item() <= 2e-5 * scale * (K ** 0.5)
    tol = 1e-5 * (K ** 0.5) if dtype == torch.bfloat16 else 1.5e-3  # tf32 truncates operands
    err = (out - ref).abs().max().item() / scale
    assert err <= tol, f"rel err {err}"


def test_gemm_epilogue_outputs():
    from cooperativeimagecaptioning_b200 import ops
    gen = torch.Generator(device="cuda").manual_seed(7)
    M, N, K = 333, 520, 256
    A = torch.randn(M, K, generator=gen, device="cuda").bfloat16()
    B = torch.randn(N, K, generator=gen, device="cuda").bfloat16()
    bias = torch.randn(N, generator=gen, device="cuda")
    rs = torch.rand(M, generator=gen, device="cuda")
    ref = torch.relu(A.float() @ B.float().t() + bias) * rs[:, None]
    out = torch.empty(M, N, device="cuda")
    out16 = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    out_t = torch.empty(N, M + 3, device="cuda", dtype=torch.bfloat16)[:, :M]
    ops.gemm(A, B, M, N, K, bias=bias, relu=True, row_scale=rs, out=out, out16=out16, out_t16=out_t)
    # accumulate mode
    acc = torch.ones(M, N, device="cuda")
    ops.gemm(A, B, M, N, K, bias=bias, relu=True, row_scale=rs, out=acc, mode=1)
    torch.cuda.synchronize()
    s = ref.abs().max().item()
    assert (out - ref).abs().max().item() <= 1e-4 * s
    assert (acc - 1 - ref).abs().max().item() <= 1e-4 * s
    assert (out16.float() - ref).abs().max().item() <= 1e-2 * s
    assert (out_t.float().t() - ref).abs().max().item() <= 1e-2 * s


def test_cast_bf16():
    from cooperativeimagecaptioning_b200 import ops
    x = torch.randn(77, 130, device="cuda")
    d = torch.empty(77, 130, device="cuda", dtype=torch.bfloat16)
    dt = torch.empty(130, 77, device="cuda", dtype=torch.bfloat16)
    ops.cast_bf16(x, d, dt)
    torch.cuda.synchronize()
    assert torch.equal(d, x.bfloat16())
    assert torch.equal(dt, x.bfloat16().t().contiguous())


@pytest.mark.parametrize("M,N,K,tile_n", [(1024, 9488, 512, 0), (333, 1000, 192, 64), (1000, 520, 256, 192)])
def test_gemm_bf16_output_and_accumulate(M, N, K, tile_n):
    """bf16 row-major output (C16), relu + bias, and the read-add-write mode (mode 1) of the
    coalesced epilogue."""
    from cooperativeimagecaptioning_b200 import ops
    gen = torch.Generator(device="cuda").manual_seed(5)
    A = torch.randn(M, K, generator=gen, device="cuda").bfloat16()
    B = torch.randn(N, K, generator=gen, device="cuda").bfloat16()
    bias = torch.randn(N, generator=gen, device="cuda")
    ref = torch.relu(A.float() @ B.float().t() + bias)
    out16 = torch.full((M, N), float("nan"), device="cuda", dtype=torch.bfloat16)
    out = torch.full((M, N), float("nan"), device="cuda")
    ops.gemm(A, B, M, N, K, bias=bias, relu=True, out=out, out16=out16, tile_n=tile_n)
    torch.cuda.synchronize()
    scale = ref.abs().max().item()
    assert (out - ref).abs().max().item() <= 2e-5 * scale * K ** 0.5
    assert (out16.float() - ref).abs().max().item() <= 1e-2 * scale
    base = torch.randn(M, N, generator=gen, device="cuda")
    acc = base.clone()
    ops.gemm(A, B, M, N, K, alpha=0.25, mode=1, out=acc, tile_n=tile_n)
    torch.cuda.synchronize()
    ref2 = base + 0.25 * (A.float() @ B.float().t())
    assert (acc - ref2).abs().max().item() <= 2e-5 * ref2.abs().max().item() * K ** 0.5
