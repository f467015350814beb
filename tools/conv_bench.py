#!/usr/bin/env python
"""Implicit-GEMM convolution kernels vs cuDNN bf16 on the same box (1 GPU), per ResNet-18/CIFAR layer shape:
forward / data gradient (native.gemm_conv) and weight gradient (native.gemm_wgrad) on operands already in bf16, against
torch's cuDNN path on bf16 channels_last tensors (aten.convolution / convolution_backward).  Our kernels additionally
keep the accumulator at (5,10) with a cast every 128 of K; cuDNN accumulates in fp32 and writes bf16.
CUDA-event timed, L2 flushed between iterations.  Prints one JSON object."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
from kernel_microbench import time_ms  # noqa: E402
from cpd_b200.ops import native  # noqa: E402
from cpd_b200.quant.functional import splits_for  # noqa: E402

# (C, H, O, k, stride) of the distinct 3x3 / 1x1 layers of ResNet-18 on 32x32 inputs
LAYERS = [(64, 32, 64, 3, 1), (64, 32, 128, 3, 2), (128, 16, 128, 3, 1), (128, 16, 256, 3, 2), (256, 8, 256, 3, 1),
          (256, 8, 512, 3, 2), (512, 4, 512, 3, 1), (64, 32, 128, 1, 2)]


def main():
    batches = [int(b) for b in sys.argv[1:]] or [64, 512]
    rows = []
    for B in batches:
        for (C, H, O, k, stride) in LAYERS:
            pad = k // 2
            OH = (H + 2 * pad - (k - 1) - 1) // stride + 1
            x = torch.randn(B, C, H, H, device="cuda").contiguous(memory_format=torch.channels_last)
            w = torch.randn(O, C, k, k, device="cuda").contiguous(memory_format=torch.channels_last)
            dy = torch.randn(B, O, OH, OH, device="cuda").contiguous(memory_format=torch.channels_last)
            xb, wb, dyb = x.bfloat16(), w.bfloat16(), dy.bfloat16()
            x_nhwc = xb.permute(0, 2, 3, 1).contiguous()
            wt = wb.permute(0, 2, 3, 1).reshape(O, k * k * C).contiguous()
            dy_rows = dyb.permute(0, 2, 3, 1).reshape(B * OH * OH, O).contiguous()
            flops = 2.0 * B * OH * OH * O * C * k * k
            fs = splits_for(B * OH * OH, O, k * k * C, 128)
            ws = splits_for(O, k * k * C, B * OH * OH, 128)
            row = {"batch": B, "C": C, "HW": H, "O": O, "k": k, "stride": stride, "GFLOP": round(flops / 1e9, 2), "fwd_splits": fs, "wgrad_splits": ws}
            ms = time_ms(lambda: native.gemm_conv(x_nhwc, wt, k, k, (OH, OH), stride, pad, 1, 5, 10, False, 128, native.OPERAND_BF16, splits=fs))
            row["ours_fwd_us"] = round(ms * 1e3, 1)
            ms = time_ms(lambda: torch.ops.aten.convolution(xb, wb, None, [stride, stride], [pad, pad], [1, 1], False, [0, 0], 1))
            row["cudnn_fwd_us"] = round(ms * 1e3, 1)
            ms = time_ms(lambda: native.gemm_wgrad(dy_rows, x_nhwc, k, k, (OH, OH), stride, pad, 1, 5, 10, False, 128, splits=ws))
            row["ours_wgrad_us"] = round(ms * 1e3, 1)
            ms = time_ms(lambda: torch.ops.aten.convolution_backward(dyb, xb, wb, None, [stride, stride], [pad, pad], [1, 1], False, [0, 0], 1,
                                                                    [False, True, False]))
            row["cudnn_wgrad_us"] = round(ms * 1e3, 1)
            if stride == 1:                                   # stride-1 data gradient = the same kernel on dy with the flipped filter
                wflip = wb.flip(2, 3).permute(1, 2, 3, 0).reshape(C, k * k * O).contiguous()
                dy_nhwc = dy_rows.view(B, OH, OH, O)
                ds = splits_for(B * H * H, C, k * k * O, 128)
                ms = time_ms(lambda: native.gemm_conv(dy_nhwc, wflip, k, k, (H, H), 1, k - 1 - pad, 1, 5, 10, False, 128, native.OPERAND_BF16, splits=ds))
                row["ours_dgrad_us"] = round(ms * 1e3, 1)
            ms = time_ms(lambda: torch.ops.aten.convolution_backward(dyb, xb, wb, None, [stride, stride], [pad, pad], [1, 1], False, [0, 0], 1,
                                                                    [True, False, False]))
            row["cudnn_dgrad_us"] = round(ms * 1e3, 1)
            rows.append(row)
            print(row, file=sys.stderr)
    print(json.dumps({"conv": rows}))


if __name__ == "__main__":
    main()
