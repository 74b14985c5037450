"""``Correlation`` / ``CorrelationFunction`` with the reference's constructor and call
signatures (correlation_package/correlation.py:6-61).  The reference's function object is a
legacy instance-style autograd.Function that current PyTorch refuses to run; here the same
surface is kept (``CorrelationFunction(pad, k, md, s1, s2, mult)(input1, input2)``) on top of a
static autograd.Function."""
from __future__ import annotations

import torch
from torch.nn.modules.module import Module

from . import correlation_cuda


class _CorrelationOp(torch.autograd.Function):
    @staticmethod
    def forward(ctx, input1, input2, pad_size, kernel_size, max_displacement, stride1, stride2, corr_multiply):
        ctx.save_for_backward(input1, input2)
        ctx.params = (pad_size, kernel_size, max_displacement, stride1, stride2, corr_multiply)
        with torch.cuda.device_of(input1):
            rbot1, rbot2, output = input1.new_empty(0), input2.new_empty(0), input1.new_empty(0)
            correlation_cuda.forward(input1, input2, rbot1, rbot2, output, *ctx.params)
        return output

    @staticmethod
    def backward(ctx, grad_output):
        input1, input2 = ctx.saved_tensors
        with torch.cuda.device_of(input1):
            rbot1, rbot2 = input1.new_empty(0), input2.new_empty(0)
            grad_input1, grad_input2 = input1.new_empty(0), input2.new_empty(0)
            correlation_cuda.backward(input1, input2, rbot1, rbot2, grad_output.contiguous(), grad_input1,
                                      grad_input2, *ctx.params)
        return grad_input1, grad_input2, None, None, None, None, None, None


class CorrelationFunction:
    def __init__(self, pad_size=3, kernel_size=3, max_displacement=20, stride1=1, stride2=2, corr_multiply=1):
        self.pad_size = pad_size
        self.kernel_size = kernel_size
        self.max_displacement = max_displacement
        self.stride1 = stride1
        self.stride2 = stride2
        self.corr_multiply = corr_multiply

    def forward(self, input1, input2):
        return _CorrelationOp.apply(input1, input2, self.pad_size, self.kernel_size, self.max_displacement,
                                    self.stride1, self.stride2, self.corr_multiply)

    __call__ = forward


class Correlation(Module):
    def __init__(self, pad_size=0, kernel_size=0, max_displacement=0, stride1=1, stride2=2, corr_multiply=1):
        super(Correlation, self).__init__()
        self.pad_size = pad_size
        self.kernel_size = kernel_size
        self.max_displacement = max_displacement
        self.stride1 = stride1
        self.stride2 = stride2
        self.corr_multiply = corr_multiply

    def forward(self, input1, input2):
        return CorrelationFunction(self.pad_size, self.kernel_size, self.max_displacement, self.stride1,
                                   self.stride2, self.corr_multiply)(input1, input2)
