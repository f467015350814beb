"""Packaging for ``pip install --no-build-isolation [-e] .`` (offline: setuptools + the in-tree build).

The native library is built by ``cpd_b200/build.py`` (nvcc -gencode arch=compute_100a,code=sm_100a, content-hashed
objects under build/, ``cpd_b200/_C*.so`` next to the package) - the same path ``python -m cpd_b200.build`` and
``__graft_entry__.build()`` use; setuptools only packages the result."""
import os
import sys

from setuptools import find_packages, setup
from setuptools.command.build_py import build_py

ROOT = os.path.dirname(os.path.abspath(__file__))


class BuildWithNative(build_py):
    def run(self):
        sys.path.insert(0, ROOT)
        from cpd_b200 import build as native_build
        native_build.build(verbose=bool(os.environ.get("CPD_B200_VERBOSE_BUILD")))
        super().run()


setup(
    name="cpd-b200",
    version="0.1.0",
    description="Blackwell-native custom-precision distributed training (capabilities and API of drcut/CPD)",
    packages=find_packages(include=["cpd_b200", "cpd_b200.*", "CPDtorch", "CPDtorch.*"]),
    package_data={"cpd_b200": ["_C*.so", "csrc/*", "csrc/*/*"]},
    python_requires=">=3.10",
    install_requires=[],            # torch, numpy and pybind11 come from the environment (no index access in the sandbox)
    cmdclass={"build_py": BuildWithNative},
)
