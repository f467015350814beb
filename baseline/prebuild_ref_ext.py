"""Pre-builds the reference's JIT extension (name 'quant_cuda', same sources, same default
flags as CPDtorch/quant/quant_function.py:9-19) into baseline/_ref/_torch_ext so the GPU box
does not spend minutes compiling at import.  Only environment is set: TORCH_EXTENSIONS_DIR
(in-tree cache), TORCH_CUDA_ARCH_LIST=10.0a (the reference passes no arch flags and there is no
GPU here to auto-detect), CXX=baseline/cxx_shim.sh (adds -DAT_CHECK=TORCH_CHECK)."""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.path.join(HERE, "_ref")


def ref_env():
    os.environ["TORCH_EXTENSIONS_DIR"] = os.path.join(REF, "_torch_ext")
    os.environ.setdefault("TORCH_CUDA_ARCH_LIST", "10.0a")
    os.environ["CXX"] = os.path.join(HERE, "cxx_shim.sh")
    os.environ.setdefault("MAX_JOBS", "4")


if __name__ == "__main__":
    ref_env()
    from torch.utils.cpp_extension import load
    cur = os.path.join(REF, "CPDtorch", "quant")
    load(name="quant_cuda",
         sources=[os.path.join(cur, "quant_cuda/quant_cuda.cpp"), os.path.join(cur, "quant_cuda/float_kernel.cu"),
                  os.path.join(cur, "quant_cuda/quant.cu")], verbose=True)
    print("reference extension built in", os.environ["TORCH_EXTENSIONS_DIR"])
