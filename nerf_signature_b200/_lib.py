"""ctypes binding of libnsig_b200.so — the thin layer between PyTorch tensors and the C ABI
declared in include/nsig.h.

There is NO fallback: if the shared library is missing or a kernel launch fails this module
raises.  Tensors are passed as raw device pointers together with torch's current CUDA stream.
"""
import ctypes
import os
import threading

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("NSIG_LIB") or os.path.join(_HERE, "libnsig_b200.so")  # NSIG_LIB: A/B builds (tools only)

_c = ctypes
_vp, _u32, _f32, _sz = _c.c_void_p, _c.c_uint32, _c.c_float, _c.c_size_t
_u64, _i64, _f64 = _c.c_uint64, _c.c_int64, _c.c_double

# name -> (argtypes, number of kernel launches one call performs)
_SIGNATURES = {
    "nsig_near_far_from_aabb": ([_vp, _vp, _vp, _u32, _f32, _vp, _vp, _vp], 1),
    "nsig_sph_from_ray": ([_vp, _vp, _f32, _u32, _vp, _vp], 1),
    "nsig_morton3D": ([_vp, _u32, _vp, _vp], 1),
    "nsig_morton3D_invert": ([_vp, _u32, _vp, _vp], 1),
    "nsig_packbits": ([_vp, _u32, _f32, _vp, _vp], 1),
    "nsig_march_rays_train": ([_vp, _vp, _vp, _f32, _f32, _u32, _u32, _u32, _u32, _u32, _vp, _vp, _vp, _vp,
                               _vp, _vp, _vp, _vp, _vp, _vp], 3),
    "nsig_march_rays_train_limited": ([_vp, _vp, _vp, _f32, _f32, _u32, _u32, _u32, _u32, _u32, _vp, _vp, _vp, _vp,
                                       _vp, _vp, _vp, _vp, _vp, _u32, _vp], 3),
    "nsig_zero_sample_padding": ([_vp, _vp, _vp, _vp, _u32, _u32, _vp], 1),
    "nsig_composite_rays_train_forward": ([_vp, _vp, _vp, _vp, _u32, _u32, _f32, _vp, _vp, _vp, _vp], 1),
    "nsig_composite_rays_train_backward": ([_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _u32, _u32, _f32, _vp, _vp,
                                            _vp], 1),
    "nsig_march_rays": ([_u32, _u32, _vp, _vp, _vp, _vp, _f32, _f32, _u32, _u32, _u32, _vp, _vp, _vp, _vp, _vp,
                         _vp, _vp, _vp], 1),
    "nsig_composite_rays_train_blend_forward": ([_vp, _vp, _vp, _vp, _u32, _u32, _f32, _f32, _vp, _vp, _vp, _vp, _vp,
                                                 _vp, _vp, _vp], 1),
    "nsig_composite_rays_train_blend_backward": ([_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _u32, _u32, _f32, _f32,
                                                  _vp, _vp, _vp], 1),
    "nsig_split_clamp_forward": ([_vp, _u32, _u32, _vp, _vp, _vp], 1),
    "nsig_split_clamp_backward": ([_vp, _vp, _vp, _u32, _u32, _vp, _vp], 1),
    "nsig_wtmk_loss_forward": ([_vp, _vp, _u32, _vp, _vp, _u32, _f32, _f32, _f32, _vp, _vp, _vp, _vp], 1),
    "nsig_wtmk_loss_backward": ([_vp, _vp, _u32, _u32, _vp, _vp, _vp, _vp], 1),
    "nsig_composite_rays": ([_u32, _u32, _f32, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp], 1),
    "nsig_hash_encode_forward": ([_vp, _u32, _vp, _vp, _u32, _u32, _vp, _vp, _vp], 1),
    "nsig_hash_encode_backward": ([_vp, _vp, _u32, _vp, _vp, _u32, _u32, _vp], 1),
    "nsig_msg_table_sum": ([_vp, _u32, _vp, _u32, _vp, _u32, _u32, _vp], 1),
    "nsig_msg_encode_forward_perbit": ([_vp, _u32, _vp, _u32, _vp, _f32, _u32, _vp, _vp], 1),
    "nsig_fused_hash_slots": ([_vp, _u32, _vp, _u32, _u32, _vp, _vp, _vp], 1),
    "nsig_field_forward": ([_vp, _vp, _u32, _f32, _vp, _vp, _u32, _vp, _f32, _vp, _vp, _f32, _vp, _vp, _vp, _vp, _vp,
                            _vp, _vp, _vp], 1),
    "nsig_field_backward_tc_masks": ([_vp, _u32, _f32, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _f32, _vp, _f32, _u32, _vp, _vp], 1),
    "nsig_field_backward_masks": ([_vp, _u32, _f32, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _f32, _vp, _f32, _u32, _vp, _vp], 1),
    "nsig_field_density": ([_vp, _u32, _f32, _vp, _vp, _u32, _vp, _f32, _vp, _f32, _vp, _vp, _vp, _vp, _vp], 1),
    "nsig_tables_to_half2": ([_vp, _u32, _u32, _vp, _vp, _vp, _vp], 2),
    "nsig_grid_sweep": ([_vp, _vp, _vp, _u32, _vp, _u64, _u32, _u32, _f64, _f32, _vp, _vp, _u32, _vp, _f32, _vp, _f32,
                         _vp, _vp, _vp, _vp], 1),
    "nsig_grid_finalize": ([_vp, _vp, _u32, _f32, _vp, _vp], 1),
    "nsig_grid_pack": ([_vp, _u32, _vp, _u32, _f32, _vp, _vp, _vp], 1),
    "nsig_grid_sample_cells": ([_vp, _u32, _u32, _u32, _u32, _u64, _vp, _vp, _vp], 4),
    "nsig_mark_untrained_grid": ([_vp, _u32, _f32, _f32, _f32, _f32, _u32, _u32, _f64, _vp, _vp], 1),
    "nsig_get_rays": ([_vp, _u32, _f32, _f32, _f32, _f32, _u32, _u32, _vp, _i64, _u32, _vp, _vp, _vp], 1),
    "nsig_allreduce_mean_inplace": ([_vp, _vp, _vp, _u32, _u32, _u32, _vp], 1),
    # kernels per call as a function of the arguments (args[4] = num_blocks): weight prep, input prep, num_blocks+1
    # convs, head / head, last-block statistics, num_blocks+1 data-gradient convs and weight-gradient kernels,
    # BatchNorm parameter gradients, input gradient
    "nsig_decoder_forward": ([_vp, _u32, _u32, _u32, _u32, _u32, _u32, _vp, _vp, _vp, _vp, _vp],
                             lambda a: a[4] + (3 if a[10] is not None else 4)),
    "nsig_decoder_backward": ([_vp, _u32, _u32, _u32, _u32, _u32, _u32, _vp, _vp, _vp, _vp, _vp, _vp],
                              # head, last block's statistics, L+1 data-gradient convs, weight gradients (the 64->64 layers as
                              # one launch per 8 layers + first + last layer), their reduction (counted here also when it is
                              # deferred to nsig_decoder_finish_backward), BatchNorm gradients (the input gradient is written
                              # by the first layer's data-gradient conv)
                              lambda a: (a[4] + 1) + 2 + (-(-max(a[4] - 1, 0) // 8) + 2) + 1 + 1),
    "nsig_decoder_prepare_weights": ([_vp, _u32, _u32, _u32, _vp, _vp], 1),
    "nsig_decoder_gelu_probe": ([_vp, _u32, _vp, _vp, _vp], 1),
    "nsig_decoder_finish_backward": ([_vp], 0),   # (its reduction launch is counted with nsig_decoder_backward)
    "nsig_color_forward": ([_vp, _vp, _u32, _vp, _vp, _vp], 1),
    "nsig_render_rays": ([_vp, _vp, _u32, _vp, _f32, _f32, _vp, _u32, _u32, _f32, _u32, _f32, _vp, _vp, _vp, _u32, _vp, _f32,
                          _vp, _vp, _f32, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp], 1),
    "nsig_msg_adam_step": ([_vp, _u32, _u32, _vp, _vp, _vp, _vp, _vp, _vp, _f32, _f32, _f32, _f32, _u32, _vp, _u32, _u32, _u32,
                            _vp], lambda a: 1 if a[17] else 2),
    "nsig_msg_adam_lookahead_sum": ([_vp, _u32, _u32, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _f32, _f32, _f32, _f32, _u32, _vp,
                                     _u32, _u32, _vp, _vp], 2),
    "nsig_msg_adam_step_sum": ([_vp, _u32, _u32, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _f32, _f32, _f32, _f32, _u32, _vp,
                                _u32, _u32, _vp, _vp], 2),
    "nsig_grad_check_update_scale": ([_vp, _u32, _vp, _vp, _f32, _f32, _c.c_int32, _vp, _vp, _vp, _vp, _vp, _vp], 1),
    "nsig_flat_adam_step": ([_vp, _vp, _vp, _vp, _u32, _vp, _vp, _vp, _f32, _vp, _f32, _f32, _f32, _vp], 1),
    "nsig_field_backward_tc": ([_vp, _vp, _u32, _f32, _vp, _vp, _vp, _vp, _vp, _f32, _vp, _f32, _u32, _vp, _vp], 1),
    "nsig_field_backward": ([_vp, _vp, _u32, _f32, _vp, _vp, _vp, _vp, _vp, _f32, _vp, _f32, _u32, _vp, _vp, _vp,
                             _vp, _vp], 1),
}

EXPORTED_SYMBOLS = sorted(list(_SIGNATURES) + ["nsig_version", "nsig_march_rays_train_scratch_bytes",
                                                "nsig_grid_sample_cells_scratch_bytes", "nsig_allreduce_grid",
                                                "nsig_decoder_workspace_bytes", "nsig_decoder_weights_bytes",
                                                "nsig_decoder_defer_weight_grads"])

_lib = None
_lock = threading.Lock()
launch_count = 0  # kernels launched through this binding (bench.py reports it as gpu_launches)


class NsigError(RuntimeError):
    pass


def load():
    """dlopen the library (once).  Raises NsigError when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            raise NsigError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(nvcc, sm_100a).  nerf_signature_b200 has no CPU or PyTorch fallback.")
        lib = ctypes.CDLL(LIB_PATH)
        for name, (argtypes, _) in _SIGNATURES.items():
            fn = getattr(lib, name)
            fn.argtypes = argtypes
            fn.restype = _c.c_int
        lib.nsig_version.restype = _c.c_char_p
        lib.nsig_version.argtypes = []
        lib.nsig_march_rays_train_scratch_bytes.restype = _sz
        lib.nsig_march_rays_train_scratch_bytes.argtypes = [_u32]
        lib.nsig_grid_sample_cells_scratch_bytes.restype = _sz
        lib.nsig_grid_sample_cells_scratch_bytes.argtypes = [_u32, _u32]
        lib.nsig_decoder_workspace_bytes.restype = _sz
        lib.nsig_decoder_workspace_bytes.argtypes = [_u32, _u32, _u32, _u32]
        lib.nsig_decoder_weights_bytes.restype = _sz
        lib.nsig_decoder_weights_bytes.argtypes = [_u32]
        lib.nsig_decoder_defer_weight_grads.restype = _c.c_int
        lib.nsig_decoder_defer_weight_grads.argtypes = [_c.c_int]
        lib.nsig_allreduce_grid.restype = _u32
        lib.nsig_allreduce_grid.argtypes = []
        _lib = lib
    return _lib


class _Ptr:
    """A device pointer argument that keeps its tensor alive until the call has been enqueued.  Call sites often pass
    temporaries (`_P(x.contiguous())`); with a bare integer the temporary would be released as soon as data_ptr()
    returns and the caching allocator could hand the same block to the next temporary of the same argument list, so
    two arguments would alias.  ctypes converts the object through `_as_parameter_`."""
    __slots__ = ("_as_parameter_", "tensor")

    def __init__(self, tensor):
        self.tensor = tensor
        self._as_parameter_ = ctypes.c_void_p(tensor.data_ptr())


def ptr(t):
    """Device pointer of a tensor (None -> NULL).  The tensor must be CUDA and contiguous; it stays referenced by the
    returned argument object for the duration of the call."""
    if t is None:
        return None
    if not t.is_cuda:
        raise NsigError("nerf_signature_b200 kernels need CUDA tensors (no CPU fallback)")
    if not t.is_contiguous():
        raise NsigError("tensor must be contiguous")
    return _Ptr(t)


def stream():
    return torch.cuda.current_stream().cuda_stream


_timing_names = ()
_timing_events = []
timing_tag = None   # label attached to the event pairs recorded from now on (a harness that captures several graphs)


def timing_enable(names):
    """Bracket every call of the named entry points with CUDA events on the launching stream
    (bench.py uses this to time the dominant kernel live inside the timed region).  Calls made while the
    stream is being captured into a CUDA graph record EXTERNAL events (event-record nodes), which are
    re-recorded by every replay: read them with timing_read() after a replay has finished."""
    global _timing_names
    _timing_names = tuple(names)
    _timing_events.clear()


def timing_reset():
    """Forget the event pairs recorded so far (keeps timing enabled)."""
    _timing_events.clear()


def timing_read(tag=None):
    """Synchronise and return {name: {"ms": total, "n": calls}} over the recorded event pairs (for graph
    replays: the most recent replay; `tag` selects the pairs recorded while `timing_tag` had that value).  Keeps the events."""
    torch.cuda.synchronize()
    out = {}
    for name, e0, e1, t in _timing_events:
        if tag is not None and t != tag:
            continue
        d = out.setdefault(name, {"ms": 0.0, "n": 0})
        d["ms"] += e0.elapsed_time(e1)
        d["n"] += 1
    return out


def timing_collect():
    """timing_read(), then disable timing and drop the events."""
    global _timing_names
    out = timing_read()
    _timing_names = ()
    _timing_events.clear()
    return out


def call(name, *args):
    """Invoke a C-ABI entry point on torch's current stream and check its status."""
    global launch_count
    lib = load()
    if name in _timing_names:
        ext = torch.cuda.is_current_stream_capturing()
        e0 = torch.cuda.Event(enable_timing=True, external=ext)
        e1 = torch.cuda.Event(enable_timing=True, external=ext)
        e0.record()
        rc = getattr(lib, name)(*args, stream())
        e1.record()
        _timing_events.append((name, e0, e1, timing_tag))
    else:
        rc = getattr(lib, name)(*args, stream())
    if rc != 0:
        if rc == -1:
            raise NsigError(f"{name}: invalid argument (NSIG_EINVAL)")
        raise NsigError(f"{name}: CUDA error {rc}")
    n = _SIGNATURES[name][1]
    launch_count += n(args) if callable(n) else n


_side_streams = {}


def side_stream(device=None, which=0):
    """A per-device auxiliary torch stream for work that is independent of the main launch chain (the message-table
    sum next to the march, the decoder's Adam next to the message-table Adam).  Fork with
    `s.wait_stream(torch.cuda.current_stream())`, join with `torch.cuda.current_stream().wait_stream(s)`; inside a
    CUDA-graph capture the pair becomes a parallel branch of the graph.  NSIG_NO_SIDE_STREAMS=1 returns None."""
    if os.environ.get("NSIG_NO_SIDE_STREAMS") == "1":
        return None
    dev = torch.cuda.current_device() if device is None else torch.device(device).index
    if dev is None:
        dev = torch.cuda.current_device()
    key = (dev, which)
    if key not in _side_streams:
        _side_streams[key] = torch.cuda.Stream(device=dev)
    return _side_streams[key]


def pointer_array(tensors):
    """Host array of device pointers (const float* const*) for a list of tensors."""
    arr = (_vp * len(tensors))(*[None if t is None else ptr(t).tensor.data_ptr() for t in tensors])
    arr._keep = list(tensors)   # the array only holds integers: keep the tensors (often temporaries) alive with it
    return arr


def float_array(values):
    return (_f32 * len(values))(*[float(v) for v in values])


def version():
    return load().nsig_version().decode()
