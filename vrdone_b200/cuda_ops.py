"""ctypes binding of libvrdone_b200.so (the C ABI declared in include/vrdone_b200.h).

``CudaOps`` is the only kernel provider of the product: there is no CPU or PyTorch fallback.  Loading fails loudly
when the shared library has not been built, and every op raises if the kernel launcher reports an error.
PyTorch is used for device memory and the current CUDA stream only.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

import torch

# VRD_LIB_PATH: A/B experiments against another build of the same ABI (tools/); the product loads the in-tree library
_LIB_PATH = os.environ.get("VRD_LIB_PATH") or os.path.join(os.path.dirname(os.path.abspath(__file__)), "libvrdone_b200.so")
_lib = None

F32, BF16 = 0, 1
_vp, _i32, _i64, _f32p = C.c_void_p, C.c_int, C.c_int64, C.c_void_p

# name -> argtypes, in the order of include/vrdone_b200.h
_SIGNATURES = {
    "vrd_abi_version": [],
    "vrd_device_arch": [],
    "vrd_set_option": [C.c_char_p, _i32],
    "vrd_get_option": [C.c_char_p],
    "vrd_h2d_pairs": [_vp, _vp, _vp, _vp, _i32, _vp],
    "vrd_merge_layout": [_i32, _vp, _vp, _vp, _vp, _vp, _vp, _vp],
    "vrd_upload": [_vp, _vp, _i64, _vp],
    "vrd_viou_filter": [_vp, _vp, _vp, _vp, _i32, C.c_float, _vp, _vp, _vp, _vp],
    "vrd_pack_pairs": [_vp, _vp, _vp, _vp, _i32, _i32, _i32, _i32, _i32, _i32, _vp, _vp, _i32, _vp, _vp, _i32, _vp],
    "vrd_gemm": [_vp, _i32, _i64, _vp, _vp, _vp, _i32, _i64, _i32, _i32, _i32, _i32, _i32, _vp, _i64, _vp, _i64, _vp, _vp,
                 _vp, _i32, _vp],
    "vrd_gemm_ln": [_vp, _i64, _vp, _vp, _vp, _vp, _vp, _i32, _vp, _i64, _i32, _i32, _i32, _i32, _vp, _vp, _i32, _vp],
    "vrd_gemm_res_ln": [_vp, _i64, _vp, _vp, _vp, _i64, _vp, _vp, _vp, _i64, _vp, _i64, _i32, _i32, _i32, _vp, _vp, _i32, _vp],
    "vrd_layernorm": [_vp, _i32, _i64, _vp, _vp, _vp, _i32, _i64, _i32, _i32, _i32, _vp, _i32, _vp],
    "vrd_small_conv": [_vp, _i32, _vp, _vp, _vp, _vp, _i32, _vp, _i32, _i64, _i32, _i32, _vp, _i32, _vp],
    "vrd_dwconv_ln": [_vp, _i32, _i64, _vp, _vp, _i32, _vp, _vp, _i32, _i32, _i32, _vp, _vp, _i32, _vp, _vp, _vp, _vp, _vp,
                      _vp, _i32, _i32, _i32, _vp],
    "vrd_window_attn": [_vp, _vp, _vp, _vp, _i32, _i64, _vp, _vp, _i32, _i32, _i32, _i32, _i32, _i32, _vp],
    "vrd_full_attn": [_vp, _vp, _vp, _vp, _i32, _i64, _vp, _vp, _i32, _i32, _i32, _i32, _i32, _vp, _i32, _vp],
    "vrd_maxpool_skip": [_vp, _i64, _vp, _vp, _i32, _vp, _vp, _i32, _i32, _vp, _i64, _i32, _vp],
    "vrd_fpn_top": [_vp, _i64, _vp, _vp, _i32, _i32, _vp, _vp, _vp, _vp, _vp, _vp, _i64, _vp],
    "vrd_fpn_level": [_vp, _i64, _vp, _i64, _vp, _vp, _i32, _vp, _vp, _i32, _i32, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i64, _vp],
    "vrd_mask_features": [_vp, _i64, _vp, _vp, _i32, _i32, _vp, _vp, _vp, _vp, _i64, _vp],
    "vrd_query_ln": [_vp, _i64, _vp, _vp, _vp, _i32, _i32, _i32, _vp, _vp, _vp, _vp, _i32, _i64, _i32, _vp],
    "vrd_query_self_attn": [_vp, _vp, _vp, _vp, _i32, _i64, _i32, _i32, _i32, _i32, _vp],
    "vrd_query_cross_attn": [_vp, _vp, _vp, _vp, _i32, _i64, _vp, _vp, _i32, _i32, _i32, _i32, _i32, _vp],
    "vrd_mask_logits": [_vp, _i64, _vp, _i64, _vp, _vp, _i32, _i32, _i32, _vp, _i64, _vp, _vp],
    "vrd_softmax_topk": [_vp, _i64, _i32, _i32, _i32, _vp, _vp, _vp],
    "vrd_rank_triplets": [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i32, _i32, _i32, _i32, _i32, _i32, _vp, _vp, _vp, _vp],
}


class ModelCfg(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("visual_dim", "clip_dim", "bbox_so_dim", "bbox_entity_dim", "embd_dim", "n_head", "fuse_head",
                                         "n_conv", "n_stem", "n_branch", "win", "use_local", "fpn_dim", "act_dtype")]


class Level(C.Structure):
    _fields_ = [("row_seq", C.c_void_p), ("seqinfo", C.c_void_p), ("R", C.c_int32), ("B", C.c_int32), ("max_len", C.c_int32),
                ("n_attn_tiles", C.c_int32), ("attn_tiles", C.c_void_p)]


_ENGINE_SYMBOLS = ["vrd_engine_last_error", "vrd_engine_create", "vrd_engine_destroy", "vrd_engine_launches",
                   "vrd_backbone_workspace_bytes", "vrd_backbone_pack", "vrd_backbone_pack_tracklets", "vrd_backbone_compute",
                   "vrd_predict_workspace_bytes", "vrd_predict"]


class PredictorCfg(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("n_embd", "num_queries", "n_head", "num_layers", "n_hidden", "n_cls", "n_cls_pad")]


def exported_symbols():
    return list(_SIGNATURES) + ["vrd_last_error"] + _ENGINE_SYMBOLS


def load_library() -> C.CDLL:
    """dlopen the in-tree shared library; raises if it is missing (build it with ``python -m vrdone_b200.build``)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(_LIB_PATH):
        raise RuntimeError(f"{_LIB_PATH} not found: the CUDA extension is not built (run `python -m vrdone_b200.build`); "
                           "vrdone_b200 has no CPU fallback")
    lib = C.CDLL(_LIB_PATH)
    for name, argtypes in _SIGNATURES.items():
        fn = getattr(lib, name)
        fn.argtypes = argtypes
        fn.restype = C.c_int
    lib.vrd_last_error.argtypes = []
    lib.vrd_last_error.restype = C.c_char_p
    lib.vrd_engine_last_error.argtypes = []
    lib.vrd_engine_last_error.restype = C.c_char_p
    lib.vrd_engine_create.argtypes = [C.POINTER(ModelCfg), C.POINTER(C.c_char_p), C.POINTER(C.c_void_p), C.POINTER(C.c_int32),
                                      C.POINTER(C.c_int32), C.c_int, C.POINTER(C.c_void_p)]
    lib.vrd_engine_create.restype = C.c_int
    lib.vrd_engine_destroy.argtypes = [C.c_void_p]
    lib.vrd_engine_destroy.restype = None
    lib.vrd_engine_launches.argtypes = [C.c_void_p]
    lib.vrd_engine_launches.restype = C.c_int64
    lib.vrd_backbone_workspace_bytes.argtypes = [C.c_void_p, C.POINTER(Level)]
    lib.vrd_backbone_workspace_bytes.restype = C.c_int64
    lib.vrd_backbone_pack.argtypes = [C.c_void_p, C.POINTER(Level), C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int64, C.c_void_p]
    lib.vrd_backbone_pack.restype = C.c_int
    lib.vrd_backbone_pack_tracklets.argtypes = [C.c_void_p, C.POINTER(Level), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_float,
                                                C.c_float, C.c_void_p, C.c_int64, C.c_void_p]
    lib.vrd_backbone_pack_tracklets.restype = C.c_int
    lib.vrd_backbone_compute.argtypes = [C.c_void_p, C.POINTER(Level), C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.vrd_backbone_compute.restype = C.c_int
    lib.vrd_predict_workspace_bytes.argtypes = [C.c_void_p, C.POINTER(PredictorCfg), C.POINTER(Level), C.POINTER(Level)]
    lib.vrd_predict_workspace_bytes.restype = C.c_int64
    lib.vrd_predict.argtypes = [C.c_void_p, C.POINTER(PredictorCfg), C.POINTER(Level), C.POINTER(Level), C.c_void_p, C.c_void_p, C.c_int,
                                C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.vrd_predict.restype = C.c_int
    if lib.vrd_abi_version() != 4:
        raise RuntimeError("libvrdone_b200.so ABI version mismatch; rebuild")
    _lib = lib
    return lib


def _dt(t: torch.Tensor) -> int:
    if t.dtype == torch.float32:
        return F32
    if t.dtype == torch.bfloat16:
        return BF16
    raise TypeError(f"unsupported dtype {t.dtype}")


def _mat(t: torch.Tensor):
    assert t.dim() == 2 and t.stride(1) == 1 and t.is_cuda, "row-major 2-D CUDA tensor expected"
    return t.data_ptr(), t.stride(0)


def _p(t: Optional[torch.Tensor]):
    if t is None:
        return None
    assert t.is_cuda and t.is_contiguous()
    return t.data_ptr()


def _f32(t: Optional[torch.Tensor]):
    if t is None:
        return None
    assert t.dtype == torch.float32
    return _p(t)


_OP_NAMES = frozenset(["pack_pairs", "gemm", "gemm_ln", "gemm_res_ln", "layernorm", "small_conv", "dwconv_ln", "window_attn", "full_attn", "maxpool_skip",
                       "fpn_top", "fpn_level", "mask_features", "query_ln", "query_self_attn", "query_cross_attn", "mask_logits",
                       "softmax_topk"])


class CudaOps:
    """Kernel provider used by ``engine.Engine``; every method enqueues one kernel on torch's current stream."""

    def __init__(self):
        self.lib = load_library()
        if not torch.cuda.is_available():
            raise RuntimeError("vrdone_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
        arch = self.lib.vrd_device_arch()
        if arch != 100:      # the binary holds sm_100a code only (arch-specific: it does not run on sm_103 / sm_120 either)
            raise RuntimeError(f"vrdone_b200 kernels are built for sm_100a only; current device is sm_{arch}")
        self.launches = 0
        self._timing = None     # list of (op name, start event, end event, algorithmic flops) while profiling
        self._last_flops = 0.0
        self._last_bytes = 0.0       # algorithmic (compulsory) HBM bytes of the last op: valid rows only, each operand once
        self._last_tag = None
        self._stream_handle = C.c_void_p(0)
        self.bind_stream()

    def set_option(self, name: str, value: int) -> int:
        """Experiment switches of the launchers ("pdl", "dw_cfg"); returns the previous value."""
        old = self.lib.vrd_set_option(name.encode(), int(value))
        if old < 0:
            raise ValueError(f"vrd_set_option: unknown option {name!r}")
        return old

    def get_option(self, name: str) -> int:
        v = self.lib.vrd_get_option(name.encode())
        if v < 0:
            raise ValueError(f"vrd_get_option: unknown option {name!r}")
        return v

    # -- per-launch CUDA-event timing (bench.py roofline pass) --------------------------------------------------------
    def start_timing(self):
        """Wrap every op of this instance with a CUDA-event pair (instance attributes shadow the class methods)."""
        self._timing = []
        for name in _OP_NAMES:
            fn = getattr(type(self), name).__get__(self)

            def timed(*a, _fn=fn, _name=name, **k):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                self._last_flops = 0.0
                self._last_bytes = 0.0
                self._last_tag = None
                e0.record()
                r = _fn(*a, **k)
                e1.record()
                # the GEMM with the LayerNorm epilogue is the same kernel: it is accounted with the GEMMs
                self._timing.append(("vrd_" + ("gemm" if _name in ("gemm_ln", "gemm_res_ln") else _name), e0, e1, self._last_flops, self._last_tag,
                                     self._last_bytes))
                return r
            setattr(self, name, timed)

    def stop_timing(self):
        torch.cuda.synchronize()
        for name in _OP_NAMES:
            if name in self.__dict__:
                delattr(self, name)
        prof = {}
        for name, e0, e1, fl, tag, nbytes in self._timing:
            ms = e0.elapsed_time(e1)
            for key in ((name,) if tag is None else (name, name + ": " + tag)):
                d = prof.setdefault(key, {"ms": 0.0, "flops": 0.0, "bytes": 0.0, "n": 0})
                d["ms"] += ms
                d["flops"] += fl
                d["bytes"] += nbytes
                d["n"] += 1
        self._timing = None
        return prof

    def bind_stream(self):
        """Cache the handle of torch's current stream for the ops that follow (one lookup per forward, not per launch)."""
        self._stream_handle = C.c_void_p(torch.cuda.current_stream().cuda_stream)

    def _stream(self):
        return self._stream_handle

    def _check(self, rc, name):
        self.launches += 1
        if rc != 0:
            raise RuntimeError(f"{name} failed: {self.lib.vrd_last_error().decode()}")

    @staticmethod
    def _lay(lay):
        return lay.row_seq.data_ptr(), lay.seqinfo.data_ptr(), lay.R

    # -- host -> device staging (copy engine; not a kernel, not counted in ``launches``) --------------------------------
    def h2d_pairs(self, src_ptrs, nbytes, dst, dst_offsets, stream):
        """src_ptrs / nbytes / dst_offsets: contiguous int64 numpy arrays; dst: CUDA tensor; stream: torch.cuda.Stream."""
        n = int(src_ptrs.shape[0])
        rc = self.lib.vrd_h2d_pairs(src_ptrs.ctypes.data, nbytes.ctypes.data, dst.data_ptr(), dst_offsets.ctypes.data, n,
                                    C.c_void_p(stream.cuda_stream))
        if rc != 0:
            raise RuntimeError(f"vrd_h2d_pairs failed: {self.lib.vrd_last_error().decode()}")

    def merge_layout(self, levels, row_seq_out, seqinfo_out):
        """levels: the LevelLayouts (one pyramid level) of <= 16 consecutive chunks, arrays on the device -> merged row_seq [sum R]
        and seqinfo [sum B, 4] written by one kernel on torch's current stream."""
        n = len(levels)
        rs = (C.c_void_p * n)(*[lv.row_seq.data_ptr() for lv in levels])
        si = (C.c_void_p * n)(*[lv.seqinfo.data_ptr() for lv in levels])
        R = (C.c_int32 * n)(*[lv.R for lv in levels])
        B = (C.c_int32 * n)(*[lv.B for lv in levels])
        st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
        self._check(self.lib.vrd_merge_layout(n, rs, si, R, B, row_seq_out.data_ptr(), seqinfo_out.data_ptr(), st), "vrd_merge_layout")

    def upload(self, host_pinned: torch.Tensor, dev: torch.Tensor, stream=None):
        """Kernel-made copy of a small PINNED host tensor into ``dev`` on ``stream`` (default: torch's current stream).  The
        caller keeps ``host_pinned`` alive until the kernel has run."""
        assert host_pinned.is_pinned() and host_pinned.is_contiguous() and dev.is_cuda and dev.is_contiguous()
        nbytes = host_pinned.numel() * host_pinned.element_size()
        assert nbytes == dev.numel() * dev.element_size()
        st = C.c_void_p((stream or torch.cuda.current_stream()).cuda_stream)
        self._check(self.lib.vrd_upload(host_pinned.data_ptr(), dev.data_ptr(), nbytes, st), "vrd_upload")

    def viou_filter(self, boxes, trk_base, durations, cat_ids, threshold, stream, want_sums=False):
        """SURVEY 8f row 2.  boxes [T, 4] fp32, trk_base / cat_ids [N] int32, durations [N, 2] int32 (CUDA tensors); enqueued on
        ``stream``.  Returns (valid [N] int32, flags [N, N] uint8, sums [N, N, 3] fp64 or None) on the device."""
        n = int(trk_base.shape[0])
        assert boxes.dtype == torch.float32 and boxes.is_contiguous() and boxes.shape[1] == 4
        assert all(t.dtype == torch.int32 and t.is_contiguous() for t in (trk_base, durations, cat_ids))
        with torch.cuda.stream(stream):
            valid = torch.empty(n, dtype=torch.int32, device=boxes.device)
            flags = torch.empty(n, n, dtype=torch.uint8, device=boxes.device)
            sums = torch.empty(n, n, 3, dtype=torch.float64, device=boxes.device) if want_sums else None
        rc = self.lib.vrd_viou_filter(_p(boxes), _p(trk_base), _p(durations), _p(cat_ids), n, float(threshold), _p(sums), _p(flags),
                                      _p(valid), C.c_void_p(stream.cuda_stream))
        self._check(rc, "vrd_viou_filter")
        return valid, flags, sums

    # -- ops ------------------------------------------------------------------------------------------------------
    def pack_pairs(self, ptrs, strides, lay, nv, nc, nbs, nbe, vis, clp, bso, bent, token_major=False):
        rs, si, R = self._lay(lay)
        frames = int(lay.len.sum())
        self._last_bytes = frames * (4.0 * (2 * nv + 2 * nc + nbs + 2 * nbe) + vis.element_size() * (2 * nv + 2 * nc) + 4.0 * 24)
        self._check(self.lib.vrd_pack_pairs(_p(ptrs), _p(strides), rs, si, R, lay.B, nv, nc, nbs, nbe, _p(vis), _p(clp),
                                            _dt(vis), _f32(bso), _f32(bent), int(token_major), self._stream()), "vrd_pack_pairs")

    def gemm(self, a, w, out, bias=None, taps=1, act=0, res1=None, res2=None, corr=None, lay=None, streams=1):
        ap, lda = _mat(a)
        op, ldo = _mat(out)
        M, K = a.shape
        N = w.shape[0]
        assert w.dtype == a.dtype and w.is_contiguous() and w.shape[1] == taps * K and out.shape == (M, N)
        r1, ld1 = _mat(res1) if res1 is not None else (None, 0)
        r2, ld2 = _mat(res2) if res2 is not None else (None, 0)
        if lay is not None:
            assert M == streams * lay.R
            rs, si, R = self._lay(lay)
        else:
            rs, si, R = None, None, 0
        valid_rows = streams * int(lay.len.sum()) if lay is not None else M
        self._last_flops = 2.0 * valid_rows * N * K * taps     # algorithmic: valid rows only (no separators / tile padding)
        self._last_bytes = valid_rows * (a.element_size() * K + out.element_size() * N + (4.0 * N if res1 is not None else 0.0)
                                         + (4.0 * N if res2 is not None else 0.0)) + w.numel() * w.element_size()
        self._last_tag = (f"{'big' if M >= 16384 else 'small'} M, {taps}x{K}->{N} {'bf16' if out.dtype == torch.bfloat16 else 'f32'}"
                          f"{' gelu' if act == 2 else ''}{' +res' if res1 is not None else ''}")
        self._check(self.lib.vrd_gemm(ap, _dt(a), lda, _p(w), _f32(bias), op, _dt(out), ldo, M, N, K, taps, act, r1, ld1, r2, ld2,
                                      _f32(corr), rs, si, R, self._stream()), "vrd_gemm")

    def gemm_ln(self, a, w, out, ln, bias=None, taps=1, corr=None, relu=False, lay=None, streams=1):
        """out (bf16) = [relu](LayerNorm_channels(a @ w.T + bias [+ corr])): the conv-as-GEMM of an embedding layer with its
        LayerNorm + ReLU as the epilogue (N = 512)."""
        ap, lda = _mat(a)
        op, ldo = _mat(out)
        M, K = a.shape
        N = w.shape[0]
        assert a.dtype == torch.bfloat16 and w.dtype == a.dtype and out.dtype == torch.bfloat16
        assert w.is_contiguous() and w.shape[1] == taps * K and out.shape == (M, N)
        if lay is not None:
            assert M == streams * lay.R
            rs, si, R = self._lay(lay)
        else:
            rs, si, R = None, None, 0
        valid_rows = streams * int(lay.len.sum()) if lay is not None else M
        self._last_flops = 2.0 * valid_rows * N * K * taps
        self._last_bytes = valid_rows * (2.0 * K + 2.0 * N) + w.numel() * 2.0
        self._last_tag = f"{'big' if M >= 16384 else 'small'} M, {taps}x{K}->{N} bf16 +ln"
        self._check(self.lib.vrd_gemm_ln(ap, lda, _p(w), _f32(bias), _f32(corr), _f32(ln[0]), _f32(ln[1]), int(relu), op, ldo, M, N, K,
                                         taps, rs, si, R, self._stream()), "vrd_gemm_ln")

    def gemm_res_ln(self, a, w, out, ln_out, ln, bias=None, res1=None, lay=None, streams=1):
        """out (fp32) = a @ w.T + bias + res1 (the residual stream) and ln_out (bf16) = LayerNorm_channels(out): the attention
        output projection of an encoder block together with the LayerNorm that feeds its MLP."""
        ap, lda = _mat(a)
        op, ldo = _mat(out)
        lp, ldl = _mat(ln_out)
        rp, ldr = _mat(res1)
        M, K = a.shape
        N = w.shape[0]
        assert a.dtype == torch.bfloat16 and w.dtype == a.dtype and out.dtype == torch.float32 and ln_out.dtype == torch.bfloat16
        assert w.is_contiguous() and w.shape[1] == K and out.shape == (M, N) and ln_out.shape == (M, N) and res1.shape == (M, N)
        if lay is not None:
            assert M == streams * lay.R
            rs, si, R = self._lay(lay)
        else:
            rs, si, R = None, None, 0
        valid_rows = streams * int(lay.len.sum()) if lay is not None else M
        self._last_flops = 2.0 * valid_rows * N * K
        self._last_bytes = valid_rows * (2.0 * K + 4.0 * N + 4.0 * N + 2.0 * N) + w.numel() * 2.0
        self._last_tag = f"{'big' if M >= 16384 else 'small'} M, 1x{K}->{N} f32 +res +ln"
        self._check(self.lib.vrd_gemm_res_ln(ap, lda, _p(w), _f32(bias), rp, ldr, _f32(ln[0]), _f32(ln[1]), op, ldo, lp, ldl, M, N, K,
                                             rs, si, R, self._stream()), "vrd_gemm_res_ln")

    def layernorm(self, x, g, b, out, relu=False, lay=None, streams=1):
        xp, ldx = _mat(x)
        op, ldo = _mat(out)
        rows, Cc = x.shape
        rs, R = (lay.row_seq.data_ptr(), lay.R) if lay is not None else (None, 1)
        valid = streams * int(lay.len.sum()) if lay is not None else rows
        self._last_bytes = valid * Cc * float(x.element_size() + out.element_size())
        self._check(self.lib.vrd_layernorm(xp, _dt(x), ldx, _f32(g), _f32(b), op, _dt(out), ldo, rows, Cc, int(relu), rs, R,
                                           self._stream()), "vrd_layernorm")

    def small_conv(self, x, cin, w, bias, ln, relu, out, lay, streams):
        op, ldo = _mat(out)
        assert x.shape[1] == 8 and x.is_contiguous() and x.shape[0] == streams * lay.R
        g, b = (ln if ln is not None else (None, None))
        self._check(self.lib.vrd_small_conv(_f32(x), cin, _f32(w), _f32(bias), _f32(g), _f32(b), int(relu), op, _dt(out), ldo,
                                            x.shape[0], w.shape[1], lay.row_seq.data_ptr(), lay.R, self._stream()),
                    "vrd_small_conv")

    def dwconv_ln(self, x, lay_in, lay_out, stride, pre, branches, streams):
        xp, ldx = _mat(x)
        n = len(branches)
        vp_arr, i32_arr, i64_arr = (C.c_void_p * n), (C.c_int32 * n), (C.c_int64 * n)
        w = vp_arr(*[_f32(br[0]) for br in branches])
        use_pre = i32_arr(*[int(br[1]) for br in branches])
        g = vp_arr(*[_f32(br[2]) for br in branches])
        b = vp_arr(*[_f32(br[3]) for br in branches])
        outs = vp_arr(*[_mat(br[4])[0] for br in branches])
        ldo = i64_arr(*[_mat(br[4])[1] for br in branches])
        odt = _dt(branches[0][4])
        assert all(_dt(br[4]) == odt for br in branches)
        pg, pb = (pre if pre is not None else (None, None))
        ri, sii, Ri = self._lay(lay_in)
        ro, sio, Ro = self._lay(lay_out)
        self._last_bytes = streams * x.shape[1] * (float(lay_in.len.sum()) * x.element_size()
                                                  + float(lay_out.len.sum()) * n * branches[0][4].element_size())
        self._check(self.lib.vrd_dwconv_ln(xp, _dt(x), ldx, ri, sii, Ri, ro, sio, Ro, lay_in.B, stride, _f32(pg), _f32(pb), n, w,
                                           use_pre, g, b, outs, ldo, odt, x.shape[1], streams, self._stream()), "vrd_dwconv_ln")

    def window_attn(self, q, k, v, out, lay, n_head, w, streams):
        qp, ld = _mat(q)
        assert k.stride(0) == ld and v.stride(0) == ld and out.stride(0) == ld
        rs, si, R = self._lay(lay)
        nl = lay.len.astype("float64")
        self._last_bytes = streams * float(nl.sum()) * q.shape[1] * 4.0 * q.element_size()
        self._last_flops = streams * 4.0 * q.shape[1] * float((nl * (2 * w + 1)).sum())      # <= 2w+1 keys per query (fewer at the edges)
        self._check(self.lib.vrd_window_attn(qp, _mat(k)[0], _mat(v)[0], _mat(out)[0], _dt(q), ld, rs, si, R, lay.B, n_head,
                                             q.shape[1], w, streams, self._stream()), "vrd_window_attn")

    def full_attn(self, q, k, v, out, lay, n_head):
        qp, ld = _mat(q)
        assert k.stride(0) == ld and v.stride(0) == ld and out.stride(0) == ld
        rs, si, R = self._lay(lay)
        nl = lay.len.astype("float64")
        self._last_flops = 4.0 * q.shape[1] * float((nl * nl).sum())        # QK^T + PV over every (query, key) pair of a pair
        self._last_bytes = float(nl.sum()) * q.shape[1] * 4.0 * q.element_size()
        tiles = getattr(lay, "tiles", None)
        self._check(self.lib.vrd_full_attn(qp, _mat(k)[0], _mat(v)[0], _mat(out)[0], _dt(q), ld, rs, si, R, lay.B, n_head,
                                           q.shape[1], lay.max_len, _p(tiles), lay.n_tiles if tiles is not None else 0,
                                           self._stream()), "vrd_full_attn")

    def maxpool_skip(self, x, lay_in, lay_out, out):
        xp, ldx = _mat(x)
        op, ldo = _mat(out)
        ri, sii, Ri = self._lay(lay_in)
        ro, sio, Ro = self._lay(lay_out)
        self._last_bytes = 4.0 * x.shape[1] * (float(lay_in.len.sum()) + float(lay_out.len.sum()))
        self._check(self.lib.vrd_maxpool_skip(xp, ldx, ri, sii, Ri, ro, sio, Ro, lay_in.B, op, ldo, x.shape[1], self._stream()),
                    "vrd_maxpool_skip")

    def fpn_top(self, x, lay, pre, w, ln, out):
        xp, ldx = _mat(x)
        op, ldo = _mat(out)
        rs, si, R = self._lay(lay)
        self._check(self.lib.vrd_fpn_top(xp, ldx, rs, si, R, lay.B, _f32(pre[0]), _f32(pre[1]), _f32(w), _f32(ln[0]), _f32(ln[1]),
                                         op, ldo, self._stream()), "vrd_fpn_top")

    def fpn_level(self, cur, y_up, lay, lay_up, ln_lat, beta_up, w, ln, out):
        cp, ldc = _mat(cur)
        up, ldu = _mat(y_up)
        op, ldo = _mat(out)
        rs, si, R = self._lay(lay)
        ru, siu, Ru = self._lay(lay_up)
        self._check(self.lib.vrd_fpn_level(cp, ldc, up, ldu, rs, si, R, ru, siu, Ru, lay.B, _f32(ln_lat[0]), _f32(ln_lat[1]),
                                           _f32(beta_up), _f32(w), _f32(ln[0]), _f32(ln[1]), op, ldo, self._stream()),
                    "vrd_fpn_level")

    def mask_features(self, y, lay, beta, w, bias, out):
        yp, ldy = _mat(y)
        op, ldo = _mat(out)
        rs, si, R = self._lay(lay)
        self._check(self.lib.vrd_mask_features(yp, ldy, rs, si, R, lay.B, _f32(beta), _f32(w), _f32(bias), op, ldo, self._stream()),
                    "vrd_mask_features")

    def query_ln(self, x, ln, pos, Q, nrows, dw, ln2, out):
        xp, ldx = _mat(x)
        op, ldo = _mat(out)
        g, b = ln if ln is not None else (None, None)
        g2, b2 = ln2 if ln2 is not None else (None, None)
        self._check(self.lib.vrd_query_ln(xp, ldx, _f32(g), _f32(b), _f32(pos), Q, nrows, out.shape[0], _f32(dw), _f32(g2), _f32(b2),
                                          op, _dt(out), ldo, x.shape[1], self._stream()), "vrd_query_ln")

    def query_self_attn(self, q, k, v, out, B, Q, n_head):
        qp, ld = _mat(q)
        assert k.stride(0) == ld and v.stride(0) == ld and out.stride(0) == ld
        self._check(self.lib.vrd_query_self_attn(qp, _mat(k)[0], _mat(v)[0], _mat(out)[0], _dt(q), ld, B, Q, n_head, q.shape[1],
                                                 self._stream()), "vrd_query_self_attn")

    def query_cross_attn(self, q, k, v, out, lay, Q, n_head):
        qp, ld = _mat(q)
        assert k.stride(0) == ld and v.stride(0) == ld and out.stride(0) == ld
        rs, si, R = self._lay(lay)
        self._check(self.lib.vrd_query_cross_attn(qp, _mat(k)[0], _mat(v)[0], _mat(out)[0], _dt(q), ld, rs, si, R, lay.B, Q, n_head,
                                                  q.shape[1], self._stream()), "vrd_query_cross_attn")

    def mask_logits(self, me, mf, lay, Q, masks, first_last):
        mp, ldm = _mat(me)
        fp, ldf = _mat(mf)
        kp, ldk = _mat(masks) if masks is not None else (None, 0)
        rs, si, R = self._lay(lay)
        assert first_last.dtype == torch.int32 and first_last.is_contiguous()
        self._check(self.lib.vrd_mask_logits(mp, ldm, fp, ldf, rs, si, R, lay.B, Q, kp, ldk, first_last.data_ptr(), self._stream()),
                    "vrd_mask_logits")

    def softmax_topk(self, logits, nrows, n_cls, topk, scores, ids):
        lp, ldl = _mat(logits)
        assert ids.dtype == torch.int32 and scores.dtype == torch.float32 and ids.is_contiguous() and scores.is_contiguous()
        self._check(self.lib.vrd_softmax_topk(lp, ldl, nrows, n_cls, topk, scores.data_ptr(), ids.data_ptr(), self._stream()),
                    "vrd_softmax_topk")

    def rank_triplets(self, scores, ids, first_last, sids, oids, cat_scores, durs, so_offset, feat_stride, pred_min_frames, n_max,
                      keys, header, records):
        """Device-side candidate filter + mean-score ranking + top-n_max cut (reference maskvrd.py:262-328); see vrd_rank_triplets."""
        B, Q, k = scores.shape
        assert scores.dtype == torch.float32 and ids.dtype == torch.int32 and first_last.dtype == torch.int32
        assert all(t.dtype == torch.int64 and t.is_cuda and t.is_contiguous() for t in (sids, oids, durs, so_offset))
        assert cat_scores.dtype == torch.float32 and keys.dtype == torch.int64 and keys.numel() >= B * Q * k
        assert header.dtype == torch.int32 and header.numel() >= 4 and records.dtype == torch.int32 and records.numel() >= 6 * n_max
        self._check(self.lib.vrd_rank_triplets(_p(scores), _p(ids), _p(first_last), _p(sids), _p(oids), _p(cat_scores), _p(durs),
                                               _p(so_offset), B, Q, k, int(feat_stride), int(pred_min_frames), int(n_max), _p(keys),
                                               _p(header), _p(records), self._stream()), "vrd_rank_triplets")
        self.launches += 1       # two kernels: keys + select


class NativeBackbone:
    """The C++ schedule of backbone + FPN (csrc/engine.cu): same kernels and order as ``engine.Engine.backbone``, issued by one
    C call per chunk.  Owns the name -> pointer table of the packed weights and a persistent activation workspace."""

    def __init__(self, ops: CudaOps, weights, mc: dict):
        self.ops, self.lib, self.weights = ops, ops.lib, weights          # ``weights`` keeps the device tensors alive
        self.device = weights.device
        clip = bool(mc.get("with_clip_feature", False))
        n_conv, n_stem, n_branch = mc["backbone_arch"]
        self.n_levels = n_branch + 1
        self.C, self.F = mc["embd_dim"], mc["fpn_dim"]
        cfg = ModelCfg(mc["visual_dim"], mc["clip_dim"] if clip else 0, mc["bbox_so_dim"], mc["bbox_entity_dim"], mc["embd_dim"],
                       mc["n_head"], mc["fuse_head"], n_conv, n_stem, n_branch, mc["n_mha_win_size"], int(bool(mc["use_local"])),
                       mc["fpn_dim"], BF16 if weights.adt == torch.bfloat16 else F32)
        names = sorted(weights.t)
        n = len(names)
        c_names = (C.c_char_p * n)(*[k.encode() for k in names])
        c_ptrs = (C.c_void_p * n)(*[weights.t[k].data_ptr() for k in names])
        c_rows = (C.c_int32 * n)(*[weights.t[k].shape[0] for k in names])
        c_cols = (C.c_int32 * n)(*[(weights.t[k].shape[1] if weights.t[k].dim() > 1 else 1) for k in names])
        handle = C.c_void_p()
        if self.lib.vrd_engine_create(C.byref(cfg), c_names, c_ptrs, c_rows, c_cols, n, C.byref(handle)) != 0:
            raise RuntimeError(f"vrd_engine_create failed: {self.lib.vrd_engine_last_error().decode()}")
        self.handle = handle
        self.workspace: Optional[torch.Tensor] = None
        self._counted = 0
        pc = mc["predictor"]
        self.Q = pc["num_queries"]
        self.pcfg = PredictorCfg(pc["n_embd"], pc["num_queries"], pc["n_head"], pc["num_layers"], pc["n_hidden"], weights.n_cls,
                                 weights.n_cls_pad)
        self.pred_workspace: Optional[torch.Tensor] = None

    def __del__(self):
        if getattr(self, "handle", None):
            self.lib.vrd_engine_destroy(self.handle)
            self.handle = None

    def _levels(self, lay):
        arr = (Level * self.n_levels)()
        for i, lv in enumerate(lay.levels):
            arr[i].row_seq, arr[i].seqinfo = lv.row_seq.data_ptr(), lv.seqinfo.data_ptr()
            arr[i].R, arr[i].B, arr[i].max_len = lv.R, lv.B, lv.max_len
            tiles = getattr(lv, "tiles", None)
            arr[i].n_attn_tiles = lv.n_tiles if tiles is not None else 0
            arr[i].attn_tiles = tiles.data_ptr() if tiles is not None and lv.n_tiles > 0 else None
        return arr

    def _fail(self, what):
        raise RuntimeError(f"{what} failed: {self.lib.vrd_engine_last_error().decode()}")

    def backbone(self, lay, pair_ptrs, pair_strides, after_pack=None, token_major=False):
        """-> (e_top [R_top, C] fp32, mask features [R_0, F] fp32), both fresh tensors."""
        def pack(lv, ws, nbytes, stream):
            if self.lib.vrd_backbone_pack(self.handle, lv, _p(pair_ptrs), _p(pair_strides), int(token_major), ws, nbytes, stream) != 0:
                self._fail("vrd_backbone_pack")
        return self._run(lay, pack, after_pack)

    def backbone_tracklets(self, lay, vis_all, clip_all, boxes_all, pair_tab, video_wh):
        """Same, with the pack stage gathering from per-tracklet arrays: vis_all [T, nv] / clip_all [T, nc] or
        None / boxes_all [T, 4] fp32 and pair_tab [B, 4] int32 on the device."""
        assert vis_all.dtype == torch.float32 and boxes_all.dtype == torch.float32 and pair_tab.dtype == torch.int32
        assert boxes_all.shape[1] == 4 and pair_tab.shape == (lay.B, 4)

        def pack(lv, ws, nbytes, stream):
            if self.lib.vrd_backbone_pack_tracklets(self.handle, lv, _p(vis_all), _p(clip_all), _p(boxes_all), _p(pair_tab),
                                                    float(video_wh[0]), float(video_wh[1]), ws, nbytes, stream) != 0:
                self._fail("vrd_backbone_pack_tracklets")
        return self._run(lay, pack, None)

    def _run(self, lay, pack, after_pack):
        lv = self._levels(lay)
        need = self.lib.vrd_backbone_workspace_bytes(self.handle, lv)
        if need < 0:
            self._fail("vrd_backbone_workspace_bytes")
        if self.workspace is None or self.workspace.numel() < need:
            self.workspace = None                                           # release before growing
            self.workspace = torch.empty(int(need * 1.1) + (1 << 20), dtype=torch.uint8, device=self.device)
        ws, nbytes = self.workspace.data_ptr(), self.workspace.numel()
        stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
        pack(lv, ws, nbytes, stream)
        if after_pack is not None:
            after_pack()
        e_top = torch.empty(lay.levels[-1].R, self.C, dtype=torch.float32, device=self.device)
        mf = torch.empty(lay.levels[0].R, self.F, dtype=torch.float32, device=self.device)
        if self.lib.vrd_backbone_compute(self.handle, lv, ws, nbytes, e_top.data_ptr(), mf.data_ptr(), stream) != 0:
            self._fail("vrd_backbone_compute")
        self._count()
        return e_top, mf

    def _count(self):
        total = int(self.lib.vrd_engine_launches(self.handle))
        self.ops.launches += total - self._counted
        self._counted = total

    def predict(self, lay, e_top, mf, topk: int, want_masks: bool = False):
        """The C++ schedule of the query decoder + heads (same kernels and order as ``engine.Engine._predictor``) over a batch
        whose levels 0 and top are described by ``lay`` (a PackLayout, or a MergedLayout over several backbone chunks)."""
        def level(lv):
            x = Level()
            x.row_seq, x.seqinfo, x.R, x.B, x.max_len = lv.row_seq.data_ptr(), lv.seqinfo.data_ptr(), lv.R, lv.B, lv.max_len
            x.n_attn_tiles, x.attn_tiles = 0, None
            return x
        l0, lt = level(lay.levels[0]), level(lay.levels[-1])
        B, Q, pc = lay.B, self.Q, self.pcfg
        need = self.lib.vrd_predict_workspace_bytes(self.handle, C.byref(pc), C.byref(l0), C.byref(lt))
        if need < 0:
            self._fail("vrd_predict_workspace_bytes")
        if self.pred_workspace is None or self.pred_workspace.numel() < need:
            self.pred_workspace = None
            self.pred_workspace = torch.empty(int(need * 1.1) + (1 << 20), dtype=torch.uint8, device=self.device)
        MQ = (B * Q + 127) // 128 * 128
        dev = self.device
        logits = torch.empty(MQ, pc.n_cls_pad, dtype=torch.float32, device=dev)
        scores = torch.empty(B * Q, topk, dtype=torch.float32, device=dev)
        ids = torch.empty(B * Q, topk, dtype=torch.int32, device=dev)
        first_last = torch.empty(B, Q, 2, dtype=torch.int32, device=dev)
        masks = torch.empty(lay.levels[0].R, Q, dtype=torch.float32, device=dev) if want_masks else None
        assert e_top.is_contiguous() and mf.is_contiguous() and e_top.dtype == torch.float32 and mf.dtype == torch.float32
        stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
        if self.lib.vrd_predict(self.handle, C.byref(pc), C.byref(l0), C.byref(lt), e_top.data_ptr(), mf.data_ptr(), int(topk),
                                self.pred_workspace.data_ptr(), self.pred_workspace.numel(), logits.data_ptr(), scores.data_ptr(),
                                ids.data_ptr(), first_last.data_ptr(), _p(masks), stream) != 0:
            self._fail("vrd_predict")
        self._count()
        return {"logits": logits[: B * Q, : pc.n_cls].view(B, Q, -1), "topk_scores": scores.view(B, Q, topk),
                "topk_ids": ids.view(B, Q, topk), "first_last": first_last, "masks": masks}
