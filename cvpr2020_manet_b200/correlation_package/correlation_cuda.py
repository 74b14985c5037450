"""Python face of the reference's pybind extension module ``correlation_cuda``
(correlation_package/correlation_cuda.cc:169-172): ``forward(...)`` and ``backward(...)`` with
the same argument lists and the same side effects on the caller's tensors (``rInput1``,
``rInput2``, ``output`` / ``gradInput*`` are resized and overwritten, cc:36-42, 107-115).
The work is done by ``manet_correlation_forward/backward`` in libmanet_b200.so."""
from __future__ import annotations

import ctypes

import torch

from .. import _lib
from .._device import check, require_cuda, stream_ptr

_DTYPES = {torch.float32: _lib.DT_F32, torch.float16: _lib.DT_F16, torch.float64: _lib.DT_F64}


def _strides(t):
    return (ctypes.c_int64 * 4)(*t.stride())


def _dtype_code(t):
    if t.dtype not in _DTYPES:
        raise TypeError(f"correlation supports float32/float16/float64 (AT_DISPATCH_FLOATING_TYPES_AND_HALF), got {t.dtype}")
    return _DTYPES[t.dtype]


def output_shape(c, h, w, pad_size, kernel_size, max_displacement, stride1, stride2):
    oc, oh, ow = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
    check(_lib.lib().manet_correlation_output_shape(c, h, w, pad_size, kernel_size, max_displacement, stride1, stride2,
                                                    ctypes.byref(oc), ctypes.byref(oh), ctypes.byref(ow)),
          "manet_correlation_output_shape")
    return oc.value, oh.value, ow.value


def _check_pair(input1, input2):
    require_cuda(input1, "input1")
    require_cuda(input2, "input2")
    if input1.dim() != 4 or input1.shape != input2.shape or input1.dtype != input2.dtype:
        raise RuntimeError("input1 and input2 must be [B,C,H,W] tensors of the same shape and dtype")


def forward(input1, input2, rInput1, rInput2, output, pad_size, kernel_size, max_displacement, stride1, stride2,
            corr_type_multiply):
    del corr_type_multiply          # accepted and unused, as in the reference kernels
    _check_pair(input1, input2)
    b, c, h, w = input1.shape
    oc, oh, ow = output_shape(c, h, w, pad_size, kernel_size, max_displacement, stride1, stride2)
    rInput1.resize_((b, h + 2 * pad_size, w + 2 * pad_size, c))
    rInput2.resize_((b, h + 2 * pad_size, w + 2 * pad_size, c))
    output.resize_((b, oc, oh, ow))
    dev = input1.device
    with torch.cuda.device(dev):
        check(_lib.lib().manet_correlation_forward(
            input1.data_ptr(), _strides(input1), input2.data_ptr(), _strides(input2), rInput1.data_ptr(),
            rInput2.data_ptr(), output.data_ptr(), b, c, h, w, pad_size, kernel_size, max_displacement, stride1,
            stride2, _dtype_code(input1), stream_ptr(dev)), "CUDA call failed (correlation forward)")
    return 1


def backward(input1, input2, rInput1, rInput2, gradOutput, gradInput1, gradInput2, pad_size, kernel_size,
             max_displacement, stride1, stride2, corr_type_multiply):
    del corr_type_multiply
    _check_pair(input1, input2)
    require_cuda(gradOutput, "gradOutput")
    b, c, h, w = input1.shape
    rInput1.resize_((b, h + 2 * pad_size, w + 2 * pad_size, c))
    rInput2.resize_((b, h + 2 * pad_size, w + 2 * pad_size, c))
    gradInput1.resize_((b, c, h, w))
    gradInput2.resize_((b, c, h, w))
    if gradOutput.dtype != input1.dtype:
        gradOutput = gradOutput.to(input1.dtype)
    dev = input1.device
    with torch.cuda.device(dev):
        check(_lib.lib().manet_correlation_backward(
            input1.data_ptr(), _strides(input1), input2.data_ptr(), _strides(input2), rInput1.data_ptr(),
            rInput2.data_ptr(), gradOutput.data_ptr(), _strides(gradOutput), gradInput1.data_ptr(),
            gradInput2.data_ptr(), b, c, h, w, pad_size, kernel_size, max_displacement, stride1, stride2,
            _dtype_code(input1), stream_ptr(dev)), "CUDA call failed (correlation backward)")
    return 1
