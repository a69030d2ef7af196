"""Kernel-level parity on the B200, through the C ABI: every entry point of include/ttvdm.h against the same op in
plain torch fp32 on identical bf16 inputs (floating-point kernels => torch fp32 reference).
Tolerance: rel-L2 < 1.5e-2 for bf16 outputs (bf16 epsilon 7.8e-3), < 1e-4 for fp32 outputs."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _cases():
    from tools import gpu_kernel_check as K
    return K.CASES, K.TOL


def pytest_generate_tests(metafunc):
    if "case_name" in metafunc.fixturenames:
        from tools import gpu_kernel_check as K
        metafunc.parametrize("case_name", [n for n, _ in K.CASES])


def test_kernel_case(case_name):
    from this_and_that_vdm_b200 import lib
    lib.init()
    cases, tol = _cases()
    fn = dict(cases)[case_name]
    n0 = lib.launch_count()
    err = fn()
    assert lib.launch_count() > n0, "no sm_100a kernel was launched"
    assert err == err and err < (1e-4 if "fp32" in case_name else tol), f"{case_name}: rel-L2 {err:.3e}"


def test_error_paths_do_not_launch():
    from this_and_that_vdm_b200 import lib
    lib.init()
    a = torch.zeros(128, 72, dtype=torch.bfloat16, device="cuda")
    w = torch.zeros(64, 72, dtype=torch.bfloat16, device="cuda")
    o = torch.zeros(128, 64, dtype=torch.bfloat16, device="cuda")
    with pytest.raises(lib.TtvdmError, match="multiples of 64"):
        lib.gemm(a, w, o, M=128, N=64, k1=72)
    with pytest.raises(lib.TtvdmError, match="empty"):
        lib.gemm(a, w, o, M=0, N=64, k1=64)
    with pytest.raises(lib.TtvdmError, match="L="):
        lib.attn_cross(a, a, a, o, ldq=64, ldo=64, rows=128, heads=1, L=200, F=1, S=128, n_ctx=1, temporal=False,
                       batch_offset=0, scale=0.125)
    with pytest.raises(lib.TtvdmError, match="F="):
        lib.attn_temporal(a, a, a, o, ldq=64, ldk=64, ldv=64, ldo=64, B=1, F=40, S=1, heads=1, scale=0.125)


def test_attention_linearity_in_v_at_full_size():
    """Size-independent property at a BASELINE-sized tile count (S = 9216 keys): attention is linear in V."""
    from this_and_that_vdm_b200 import lib
    lib.init()
    n, heads, S = 1, 2, 9216
    C = heads * 64
    g = torch.Generator(device="cpu").manual_seed(5)
    q = torch.randn(n * S, C, generator=g).to("cuda", torch.bfloat16)
    k = torch.randn(n * S, C, generator=g).to("cuda", torch.bfloat16)
    v1 = torch.randn(n * S, C, generator=g).to("cuda", torch.bfloat16)
    v2 = torch.randn(n * S, C, generator=g).to("cuda", torch.bfloat16)
    outs = []
    for v in (v1, v2, (v1.float() + v2.float()).to(torch.bfloat16)):
        o = torch.empty(n * S, C, dtype=torch.bfloat16, device="cuda")
        lib.attn_spatial(q, k, v, o, ldq=C, ldk=C, ldv=C, ldo=C, n_img=n, heads=heads, seq=S, scale=0.125)
        outs.append(o.float())
    torch.cuda.synchronize()
    err = float((outs[0] + outs[1] - outs[2]).norm() / outs[2].norm())
    assert err < 2e-2, err
    # rows of softmax sum to one: V = const => output = const
    ones = torch.full((n * S, C), 0.5, dtype=torch.bfloat16, device="cuda")
    o = torch.empty(n * S, C, dtype=torch.bfloat16, device="cuda")
    lib.attn_spatial(q, k, ones, o, ldq=C, ldk=C, ldv=C, ldo=C, n_img=n, heads=heads, seq=S, scale=0.125)
    assert float((o.float() - 0.5).abs().max()) < 5e-3
