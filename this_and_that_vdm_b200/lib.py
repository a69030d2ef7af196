"""ctypes binding of libttvdm_sm100.so (C ABI in include/ttvdm.h).

Only raw device pointers, sizes and the CUDA stream cross this boundary; torch is used purely as the owner of
device memory. There is no CPU fallback: if the shared library is missing or the device is not sm_100 every
call raises.
"""
from __future__ import annotations

import ctypes as C
from pathlib import Path
from typing import Optional

import torch

LIB_PATH = Path(__file__).resolve().parent / "libttvdm_sm100.so"

A_LINEAR, A_CONV3X3, A_TCONV3 = 0, 1, 2

c_void_p, c_int, c_float, c_size_t = C.c_void_p, C.c_int, C.c_float, C.c_size_t


class GemmParams(C.Structure):
    _fields_ = [
        ("mode", c_int), ("a", c_void_p), ("a2", c_void_p), ("k1", c_int), ("k2", c_int),
        ("lda", c_int), ("lda2", c_int), ("n_img", c_int), ("H", c_int), ("W", c_int),
        ("w", c_void_p), ("M", c_int), ("N", c_int),
        ("bias", c_void_p), ("rowvec", c_void_p), ("rows_per_vec", c_int), ("ldrv", c_int), ("s0", c_float),
        ("res1", c_void_p), ("ldr1", c_int), ("s1", c_float),
        ("res2", c_void_p), ("ldr2", c_int), ("s2", c_float),
        ("geglu", c_int), ("out", c_void_p), ("ldo", c_int), ("out_fp32", c_int), ("act", c_int),
        # normalisation fusions (ABI 3)
        ("gn_stats_out", c_void_p), ("gn_rows_per_inst", c_int), ("row_sums_out", c_void_p),
        ("rs_addvec", c_void_p), ("rs_add_rows", c_int), ("rs_add_mod", c_int), ("ld_rs_add", c_int),
        ("ln_rowsums", c_void_p), ("ln_colsum", c_void_p), ("ln_eps", c_float),
        ("prevec", c_void_p), ("prevec_rows", c_int), ("prevec_mod", c_int), ("ldpv", c_int),
        ("ln_row_add", c_void_p), ("conv_stride", c_int), ("conv_taps", c_int), ("conv_dy0", c_int),
        ("conv_dx0", c_int),
    ]


class AttnParams(C.Structure):
    _fields_ = [
        ("q", c_void_p), ("k", c_void_p), ("v", c_void_p), ("out", c_void_p),
        ("ldq", c_int), ("ldk", c_int), ("ldv", c_int), ("ldo", c_int),
        ("n_img", c_int), ("heads", c_int), ("seq", c_int), ("scale", c_float),
    ]


class XAttnParams(C.Structure):
    _fields_ = [
        ("q", c_void_p), ("kc", c_void_p), ("vc", c_void_p), ("out", c_void_p),
        ("ldq", c_int), ("ldo", c_int), ("rows", c_int), ("heads", c_int), ("L", c_int),
        ("F", c_int), ("S", c_int), ("n_ctx", c_int), ("temporal", c_int), ("batch_offset", c_int),
        ("scale", c_float), ("head_dim", c_int),
    ]


class TAttnParams(C.Structure):
    _fields_ = [
        ("q", c_void_p), ("k", c_void_p), ("v", c_void_p), ("out", c_void_p),
        ("ldq", c_int), ("ldk", c_int), ("ldv", c_int), ("ldo", c_int),
        ("B", c_int), ("F", c_int), ("S", c_int), ("heads", c_int), ("scale", c_float), ("head_dim", c_int),
    ]


class GroupNormParams(C.Structure):
    _fields_ = [
        ("x1", c_void_p), ("c1", c_int), ("ld1", c_int),
        ("x2", c_void_p), ("c2", c_int), ("ld2", c_int),
        ("rows", c_int), ("rows_per_inst", c_int), ("eps", c_float),
        ("stats", c_void_p), ("gamma", c_void_p), ("beta", c_void_p), ("silu", c_int),
        ("out", c_void_p), ("ldo", c_int),
        ("pstats1", c_void_p), ("pstats2", c_void_p),
    ]


class LayerNormParams(C.Structure):
    _fields_ = [
        ("x", c_void_p), ("ldx", c_int), ("rows", c_int), ("C", c_int),
        ("addvec", c_void_p), ("F", c_int), ("S", c_int),
        ("sum_out", c_void_p), ("ldsum", c_int),
        ("gamma", c_void_p), ("beta", c_void_p), ("eps", c_float),
        ("out", c_void_p), ("ldo", c_int),
    ]


class PrepareParams(C.Structure):
    _fields_ = [
        ("latents", c_void_p), ("image_latents", c_void_p), ("cond", c_void_p),
        ("model_in", c_void_p), ("c_pad", c_int),
        ("B_local", c_int), ("batch_offset", c_int), ("F", c_int), ("h", c_int), ("w", c_int),
        ("sigma", c_float),
    ]


class EulerParams(C.Structure):
    _fields_ = [
        ("latents", c_void_p), ("eps_u", c_void_p), ("eps_c", c_void_p), ("ld_eps", c_int),
        ("guidance", c_void_p), ("F", c_int), ("h", c_int), ("w", c_int),
        ("sigma", c_float), ("sigma_next", c_float),
    ]


GESTURE_MAX_POINTS = 16


class GestureParams(C.Structure):
    _fields_ = [
        ("n_points", c_int),
        ("frame_idx", c_int * GESTURE_MAX_POINTS), ("vertical", c_int * GESTURE_MAX_POINTS),
        ("horizontal", c_int * GESTURE_MAX_POINTS),
        ("org_h", c_int), ("org_w", c_int), ("H", c_int), ("W", c_int), ("F", c_int),
        ("dilate", c_int), ("flip", c_int),
        ("scratch", c_void_p), ("out", c_void_p),
    ]


class PackLinearParams(C.Structure):
    _fields_ = [
        ("w", c_void_p), ("bias", c_void_p), ("gamma", c_void_p), ("beta", c_void_p), ("src_dtype", c_int),
        ("N", c_int), ("K", c_int), ("geglu", c_int), ("out_row0", c_int), ("ldo", c_int),
        ("out_w", c_void_p), ("out_bias", c_void_p), ("out_colsum", c_void_p),
    ]


EXPORTS = [
    "ttvdm_init", "ttvdm_destroy", "ttvdm_last_error", "ttvdm_abi_version", "ttvdm_launch_count",
    "ttvdm_pack_conv_weight", "ttvdm_pack_linear", "ttvdm_pack_vector",
    "ttvdm_groupnorm_workspace_bytes", "ttvdm_gemm_gn_stats_bytes", "ttvdm_gemm_row_sums_bytes",
    "ttvdm_gemm_workspace_bytes", "ttvdm_attn_workspace_bytes", "ttvdm_gesture_scratch_bytes",
    "ttvdm_gemm", "ttvdm_attn_spatial", "ttvdm_attn_cross", "ttvdm_attn_temporal",
    "ttvdm_groupnorm", "ttvdm_layernorm", "ttvdm_im2col_s2", "ttvdm_upsample2x", "ttvdm_interleave2x", "ttvdm_axpy", "ttvdm_sinusoid",
    "ttvdm_sampler_prepare", "ttvdm_sampler_euler_step", "ttvdm_gesture_raster",
    "ttvdm_softmax_rows", "ttvdm_im2col_s2_pad01", "ttvdm_vae_time_conv_out",
    "ttvdm_act_inplace", "ttvdm_layernorm_flat",
]

_lib: Optional[C.CDLL] = None
_inited_devices: set[int] = set()


class TtvdmError(RuntimeError):
    pass


def load() -> C.CDLL:
    """dlopen the in-tree library (no compute). Raises if it has not been built."""
    global _lib
    if _lib is None:
        if not LIB_PATH.exists():
            raise TtvdmError(
                f"{LIB_PATH} is missing: run `python -m this_and_that_vdm_b200.build` (there is no CPU fallback)")
        lib = C.CDLL(str(LIB_PATH))
        lib.ttvdm_launch_count.restype = C.c_uint64
        for q in ("ttvdm_groupnorm_workspace_bytes", "ttvdm_gemm_gn_stats_bytes", "ttvdm_gemm_row_sums_bytes",
                  "ttvdm_gemm_workspace_bytes", "ttvdm_attn_workspace_bytes", "ttvdm_gesture_scratch_bytes"):
            getattr(lib, q).restype = C.c_size_t
        for name in EXPORTS:
            getattr(lib, name)  # AttributeError if a declared symbol is not exported
        _lib = lib
    return _lib


def _check(rc: int, what: str) -> None:
    if rc != 0:
        buf = C.create_string_buffer(512)
        load().ttvdm_last_error(buf, 512)
        raise TtvdmError(f"{what} failed (status {rc}): {buf.value.decode(errors='replace')}")


def init(device: Optional[int] = None) -> None:
    if not torch.cuda.is_available():
        raise TtvdmError("ttvdm needs a CUDA device (sm_100); there is no CPU fallback")
    dev = torch.cuda.current_device() if device is None else device
    if dev not in _inited_devices:
        _check(load().ttvdm_init(dev), "ttvdm_init")
        _inited_devices.add(dev)


_graph_launches = 0


def launch_count() -> int:
    """Kernels of this library enqueued so far: direct launches (counted inside the library) plus the kernel nodes of
    every CUDA-graph replay reported through add_graph_launches()."""
    return int(load().ttvdm_launch_count()) + _graph_launches


def add_graph_launches(n: int) -> None:
    global _graph_launches
    _graph_launches += int(n)


def destroy() -> None:
    """ttvdm_destroy: forget the process-wide state; the next init() / call initialises again."""
    _check(load().ttvdm_destroy(), "ttvdm_destroy")
    _inited_devices.clear()


def _stream() -> c_void_p:
    return c_void_p(torch.cuda.current_stream().cuda_stream)


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


# Optional per-call CUDA-event timing (bench.py's roofline pass): records (entry point, params, start, end).
_profile: Optional[list] = None


def profiling() -> bool:
    """True while start_profile() is active (per-call CUDA events are not graph capturable: callers stay eager)."""
    return _profile is not None


def start_profile() -> None:
    global _profile
    _profile = []


def stop_profile() -> list:
    """Returns [(name, info dict, milliseconds)] for every call since start_profile(); synchronises once."""
    global _profile
    rec, _profile = _profile or [], None
    torch.cuda.synchronize()
    return [(n, info, e0.elapsed_time(e1)) for (n, info, e0, e1) in rec]


def _info(name: str, p) -> dict:
    if name == "ttvdm_gemm":
        taps = 9 if p.mode == A_CONV3X3 else (3 if p.mode == A_TCONV3 else 1)
        return {"flops": 2.0 * p.M * p.N * taps * (p.k1 + (p.k2 if p.a2 else 0)), "mode": p.mode, "M": p.M, "N": p.N,
                "K": taps * (p.k1 + (p.k2 if p.a2 else 0))}
    if name == "ttvdm_attn_spatial":
        return {"flops": 4.0 * p.n_img * p.heads * p.seq * p.seq * 64}
    if name == "ttvdm_attn_cross":
        return {"flops": 4.0 * p.rows * p.heads * p.L * 64}
    if name == "ttvdm_attn_temporal":
        return {"flops": 4.0 * p.B * p.S * p.heads * p.F * p.F * 64}
    return {"flops": 0.0}


def call(name: str, params: C.Structure) -> None:
    if _profile is None:
        _check(getattr(load(), name)(C.byref(params), _stream()), name)
        return
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    _check(getattr(load(), name)(C.byref(params), _stream()), name)
    e1.record()
    _profile.append((name, _info(name, params), e0, e1))


def call_raw(name: str, *args) -> None:
    if _profile is None:
        _check(getattr(load(), name)(*args, _stream()), name)
        return
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    _check(getattr(load(), name)(*args, _stream()), name)
    e1.record()
    _profile.append((name, {"flops": 0.0}, e0, e1))


# ------------------------------------------------------------------------------------------------ wrappers
def gemm(a: torch.Tensor, w: torch.Tensor, out: torch.Tensor, *, M: int, N: int, k1: int,
         mode: int = A_LINEAR, lda: Optional[int] = None, a2: Optional[torch.Tensor] = None, k2: int = 0,
         lda2: int = 0, n_img: int = 0, H: int = 0, W: int = 0, bias: Optional[torch.Tensor] = None,
         rowvec: Optional[torch.Tensor] = None, rows_per_vec: int = 0, ldrv: int = 0, s0: float = 1.0,
         res1: Optional[torch.Tensor] = None, ldr1: int = 0, s1: float = 1.0,
         res2: Optional[torch.Tensor] = None, ldr2: int = 0, s2: float = 1.0,
         geglu: bool = False, ldo: Optional[int] = None, out_fp32: bool = False, act: int = 0,
         gn_stats_out: Optional[torch.Tensor] = None, gn_rows_per_inst: int = 0,
         row_sums_out: Optional[torch.Tensor] = None, rs_addvec: Optional[torch.Tensor] = None, rs_add_rows: int = 0,
         rs_add_mod: int = 0,
         ln_rowsums: Optional[torch.Tensor] = None, ln_colsum: Optional[torch.Tensor] = None, ln_eps: float = 1e-5,
         prevec: Optional[torch.Tensor] = None, prevec_rows: int = 0, prevec_mod: int = 0, ldpv: int = 0,
         ln_row_add: Optional[torch.Tensor] = None, conv_stride: int = 1, conv_taps: int = 0, conv_dy0: int = 0,
         conv_dx0: int = 0) -> None:
    """gn_stats_out: fp64 [M / gn_rows_per_inst, N / 2, 2] (pre-zeroed; the epilogue adds the GroupNorm sums of `out`);
    row_sums_out: fp32 [N / 32, M, 2] (LayerNorm partial sums of `out`, no zeroing needed); ln_rowsums (= the producer's
    row_sums_out, [K / 32, M, 2]) / ln_colsum (+ prevec / ln_row_add):
    LayerNorm of the A operand folded into this GEMM's epilogue — see include/ttvdm.h."""
    p = GemmParams()
    p.mode = mode
    p.a, p.a2, p.k1, p.k2 = _ptr(a), _ptr(a2), k1, k2
    p.lda = k1 if lda is None else lda
    p.lda2 = lda2 if lda2 else k2
    p.n_img, p.H, p.W = n_img, H, W
    p.w, p.M, p.N = _ptr(w), M, N
    p.bias, p.rowvec, p.rows_per_vec, p.ldrv, p.s0 = _ptr(bias), _ptr(rowvec), rows_per_vec, ldrv, s0
    p.res1, p.ldr1, p.s1 = _ptr(res1), ldr1 if ldr1 else N, s1
    p.res2, p.ldr2, p.s2 = _ptr(res2), ldr2 if ldr2 else N, s2
    p.geglu = int(geglu)
    p.out = _ptr(out)
    p.ldo = ldo if ldo is not None else (N // 2 if geglu else N)
    p.out_fp32 = int(out_fp32)
    p.act = act
    p.gn_stats_out, p.gn_rows_per_inst, p.row_sums_out = _ptr(gn_stats_out), gn_rows_per_inst, _ptr(row_sums_out)
    p.rs_addvec, p.rs_add_rows, p.rs_add_mod, p.ld_rs_add = _ptr(rs_addvec), rs_add_rows, rs_add_mod, N
    p.ln_rowsums, p.ln_colsum, p.ln_eps = _ptr(ln_rowsums), _ptr(ln_colsum), ln_eps
    p.prevec, p.prevec_rows, p.prevec_mod, p.ldpv = _ptr(prevec), prevec_rows, prevec_mod, ldpv if ldpv else N
    p.ln_row_add = _ptr(ln_row_add)
    p.conv_stride, p.conv_taps, p.conv_dy0, p.conv_dx0 = conv_stride, conv_taps, conv_dy0, conv_dx0
    call("ttvdm_gemm", p)


def attn_spatial(q, k, v, out, *, ldq, ldk, ldv, ldo, n_img, heads, seq, scale) -> None:
    p = AttnParams()
    p.q, p.k, p.v, p.out = _ptr(q), _ptr(k), _ptr(v), _ptr(out)
    p.ldq, p.ldk, p.ldv, p.ldo = ldq, ldk, ldv, ldo
    p.n_img, p.heads, p.seq, p.scale = n_img, heads, seq, scale
    call("ttvdm_attn_spatial", p)


def attn_cross(q, kc, vc, out, *, ldq, ldo, rows, heads, L, F, S, n_ctx, temporal, batch_offset, scale,
               head_dim=64) -> None:
    p = XAttnParams()
    p.q, p.kc, p.vc, p.out = _ptr(q), _ptr(kc), _ptr(vc), _ptr(out)
    p.ldq, p.ldo, p.rows, p.heads, p.L = ldq, ldo, rows, heads, L
    p.F, p.S, p.n_ctx, p.temporal, p.batch_offset, p.scale = F, S, n_ctx, int(temporal), batch_offset, scale
    p.head_dim = head_dim
    call("ttvdm_attn_cross", p)


def attn_temporal(q, k, v, out, *, ldq, ldk, ldv, ldo, B, F, S, heads, scale, head_dim=64) -> None:
    p = TAttnParams()
    p.q, p.k, p.v, p.out = _ptr(q), _ptr(k), _ptr(v), _ptr(out)
    p.ldq, p.ldk, p.ldv, p.ldo = ldq, ldk, ldv, ldo
    p.B, p.F, p.S, p.heads, p.scale, p.head_dim = B, F, S, heads, scale, head_dim
    call("ttvdm_attn_temporal", p)


def groupnorm(x1, out, stats, gamma, beta, *, c1, rows, rows_per_inst, eps, silu, x2=None, c2=0,
              ld1=None, ld2=None, ldo=None, pstats1=None, pstats2=None) -> None:
    """pstats1 / pstats2: fp64 [n_inst, c / 2, 2] GroupNorm sums written by the epilogue of the GEMM that produced x1 / x2
    (gemm(gn_stats_out=...)); a source that has them is not read by the statistics pass (stats may then be None)."""
    p = GroupNormParams()
    p.x1, p.c1, p.ld1 = _ptr(x1), c1, c1 if ld1 is None else ld1
    p.x2, p.c2, p.ld2 = _ptr(x2), c2, c2 if ld2 is None else ld2
    p.rows, p.rows_per_inst, p.eps = rows, rows_per_inst, eps
    p.stats, p.gamma, p.beta, p.silu = _ptr(stats), _ptr(gamma), _ptr(beta), int(silu)
    p.out, p.ldo = _ptr(out), (c1 + c2) if ldo is None else ldo
    p.pstats1, p.pstats2 = _ptr(pstats1), _ptr(pstats2)
    call("ttvdm_groupnorm", p)


def layernorm(x, out, gamma, beta, *, rows, C, eps=1e-5, addvec=None, F=0, S=0, sum_out=None,
              ldx=None, ldo=None, ldsum=None) -> None:
    p = LayerNormParams()
    p.x, p.ldx, p.rows, p.C = _ptr(x), C if ldx is None else ldx, rows, C
    p.addvec, p.F, p.S = _ptr(addvec), F, S
    p.sum_out, p.ldsum = _ptr(sum_out), C if ldsum is None else ldsum
    p.gamma, p.beta, p.eps = _ptr(gamma), _ptr(beta), eps
    p.out, p.ldo = _ptr(out), C if ldo is None else ldo
    call("ttvdm_layernorm", p)


def interleave2x(parts, out, *, n_img, H, W, C) -> None:
    call_raw("ttvdm_interleave2x", c_void_p(_ptr(parts)), c_void_p(_ptr(out)), n_img, H, W, C)


def im2col_s2(x, out, *, n_img, H, W, C) -> None:
    call_raw("ttvdm_im2col_s2", c_void_p(_ptr(x)), c_void_p(_ptr(out)), n_img, H, W, C)


def upsample2x(x, out, *, n_img, H, W, C) -> None:
    call_raw("ttvdm_upsample2x", c_void_p(_ptr(x)), c_void_p(_ptr(out)), n_img, H, W, C)


def sinusoid(t, out, *, n: int, dim: int) -> None:
    call_raw("ttvdm_sinusoid", c_void_p(_ptr(t)), c_void_p(_ptr(out)), n, dim)


def axpy(a, b, out, scale: float, n: int) -> None:
    call_raw("ttvdm_axpy", c_void_p(_ptr(a)), c_void_p(_ptr(b)), c_void_p(_ptr(out)), c_float(scale), c_size_t(n))


def sampler_prepare(latents, image_latents, cond, model_in, *, c_pad, B_local, batch_offset, F, h, w, sigma) -> None:
    p = PrepareParams()
    p.latents, p.image_latents, p.cond = _ptr(latents), _ptr(image_latents), _ptr(cond)
    p.model_in, p.c_pad = _ptr(model_in), c_pad
    p.B_local, p.batch_offset, p.F, p.h, p.w, p.sigma = B_local, batch_offset, F, h, w, sigma
    call("ttvdm_sampler_prepare", p)


def gesture_raster(points, out, scratch, *, org_h, org_w, dilate=True, flip=False) -> None:
    """points: sequence of (frame_idx, vertical, horizontal) in data.txt order; out: fp32 [F, 3, H, W] on the device;
    scratch: fp32 device buffer with >= len(points) * (H + W) elements."""
    if out.dtype != torch.float32 or out.dim() != 4 or out.shape[1] != 3 or not out.is_contiguous():
        raise TtvdmError("gesture_raster: out must be a contiguous fp32 [F, 3, H, W] tensor")
    if len(points) > GESTURE_MAX_POINTS:
        raise TtvdmError(f"gesture_raster: at most {GESTURE_MAX_POINTS} points")
    F, _, H, W = out.shape
    if scratch.dtype != torch.float32 or scratch.numel() < max(1, len(points)) * (H + W):
        raise TtvdmError("gesture_raster: scratch must hold len(points) * (H + W) floats")
    p = GestureParams()
    p.n_points = len(points)
    for i, (f, v, h) in enumerate(points):
        p.frame_idx[i], p.vertical[i], p.horizontal[i] = int(f), int(v), int(h)
    p.org_h, p.org_w, p.H, p.W, p.F = int(org_h), int(org_w), H, W, F
    p.dilate, p.flip = int(bool(dilate)), int(bool(flip))
    p.scratch, p.out = _ptr(scratch), _ptr(out)
    call("ttvdm_gesture_raster", p)


def sampler_euler_step(latents, eps_u, eps_c, guidance, *, ld_eps, F, h, w, sigma, sigma_next) -> None:
    p = EulerParams()
    p.latents, p.eps_u, p.eps_c, p.ld_eps = _ptr(latents), _ptr(eps_u), _ptr(eps_c), ld_eps
    p.guidance, p.F, p.h, p.w, p.sigma, p.sigma_next = _ptr(guidance), F, h, w, sigma, sigma_next
    call("ttvdm_sampler_euler_step", p)


# ---- weight repack entry points (include/ttvdm.h, "Weight repack entry points"): once per model load
_DT = {torch.float32: 0, torch.float16: 1, torch.bfloat16: 2}


def _src(t: torch.Tensor) -> int:
    if t.dtype not in _DT or not t.is_contiguous():
        raise TtvdmError(f"pack: source must be a contiguous fp32 / fp16 / bf16 tensor (got {t.dtype})")
    return _DT[t.dtype]


def pack_conv_weight(w: torch.Tensor, out: torch.Tensor, *, cin_pad: int = 0) -> None:
    """w [cout, cin, *taps] (3x3 -> 9 taps, (3,1,1) -> 3, 1x1 -> 1) -> out bf16 [cout, taps * cin_pad] tap-major."""
    cout, cin = w.shape[:2]
    taps = w[0, 0].numel()
    call_raw("ttvdm_pack_conv_weight", c_void_p(_ptr(w)), _src(w), cout, cin, taps, max(cin_pad, cin), c_void_p(_ptr(out)))


def pack_linear(w: torch.Tensor, out_w: torch.Tensor, *, bias=None, gamma=None, beta=None, out_bias=None, out_colsum=None,
                geglu: bool = False, out_row0: int = 0) -> None:
    """out_w[row] = bf16(w[n] * gamma); out_colsum[row] = sum_k out_w[row, k]; out_bias[row] = bias[n] + w[n] . beta."""
    p = PackLinearParams()
    N = w.shape[0]
    K = w[0].numel()
    for t in (bias, gamma, beta):
        if t is not None and (t.dtype != w.dtype or not t.is_contiguous()):
            raise TtvdmError("pack_linear: bias / gamma / beta must share the weight's dtype")
    p.w, p.bias, p.gamma, p.beta, p.src_dtype = _ptr(w), _ptr(bias), _ptr(gamma), _ptr(beta), _src(w)
    p.N, p.K, p.geglu, p.out_row0, p.ldo = N, K, int(geglu), out_row0, out_w.shape[-1]
    p.out_w, p.out_bias, p.out_colsum = _ptr(out_w), _ptr(out_bias), _ptr(out_colsum)
    call("ttvdm_pack_linear", p)


def pack_vector(src: torch.Tensor, out: torch.Tensor) -> None:
    call_raw("ttvdm_pack_vector", c_void_p(_ptr(src)), _src(src), c_size_t(src.numel()), c_void_p(_ptr(out)))


# ---- VAE-only entry points (include/ttvdm.h, "VAE either side of the loop")
def softmax_rows(x, out, *, rows, cols, ldx, ldo, cols_out, causal=False) -> None:
    """x fp32 [rows, cols] (ldx) -> out bf16 [rows, cols_out] (ldo); columns >= cols are written as zeros. causal: False /
    0 = no mask; an int P = rows come in blocks of P queries, row r attends to columns 0..r % P; True = one block (P = rows)."""
    causal = rows if causal is True else int(causal)
    if x.dtype != torch.float32 or out.dtype != torch.bfloat16:
        raise TtvdmError("softmax_rows: x must be fp32 and out bf16")
    call_raw("ttvdm_softmax_rows", c_void_p(_ptr(x)), ldx, c_void_p(_ptr(out)), ldo, rows, cols, cols_out, causal)


def im2col_s2_pad01(x, out, *, n_img, H, W, C) -> None:
    call_raw("ttvdm_im2col_s2_pad01", c_void_p(_ptr(x)), c_void_p(_ptr(out)), n_img, H, W, C)


def vae_time_conv_out(x, w, bias, out, *, B, F, H, W, ldx) -> None:
    """x: device fp32 [(b, f, s), ldx]; w: HOST fp32 [3, 3, 3] (co, ci, t); bias: HOST fp32 [3];
    out: device fp32 [B*F, 3, H, W]."""
    if w.is_cuda or bias.is_cuda or w.dtype != torch.float32 or bias.dtype != torch.float32 or w.numel() != 27:
        raise TtvdmError("vae_time_conv_out: w / bias must be host fp32 tensors with 27 / 3 elements")
    if x.dtype != torch.float32 or out.dtype != torch.float32 or not out.is_contiguous():
        raise TtvdmError("vae_time_conv_out: x / out must be fp32 (out contiguous)")
    w, bias = w.contiguous(), bias.contiguous()
    call_raw("ttvdm_vae_time_conv_out", c_void_p(_ptr(x)), ldx, c_void_p(w.data_ptr()), c_void_p(bias.data_ptr()),
             c_void_p(_ptr(out)), B, F, H * W)


# ---- conditioning-builder entry points (include/ttvdm.h, "Conditioning builder")
ACT_GELU, ACT_QUICK_GELU = 2, 3


def act_inplace(x, kind: int) -> None:
    if x.dtype != torch.bfloat16 or not x.is_contiguous():
        raise TtvdmError("act_inplace: x must be a contiguous bf16 tensor")
    call_raw("ttvdm_act_inplace", c_void_p(_ptr(x)), c_size_t(x.numel()), kind)


def layernorm_flat(x, out, *, rows, n, eps=1e-5) -> None:
    if x.dtype != torch.float32 or out.dtype != torch.float32 or not x.is_contiguous() or not out.is_contiguous():
        raise TtvdmError("layernorm_flat: x / out must be contiguous fp32 tensors")
    call_raw("ttvdm_layernorm_flat", c_void_p(_ptr(x)), c_void_p(_ptr(out)), rows, c_size_t(n), c_float(eps))
