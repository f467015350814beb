#!/bin/bash
# Host-compiler shim for building the reference's JIT extension on torch >= 2.x:
# CPDtorch/quant/quant_cuda/quant_cuda.cpp:8,10 uses AT_CHECK, which PyTorch removed.
# The macro is supplied on the command line; no reference file is edited.
exec g++ -DAT_CHECK=TORCH_CHECK "$@"
