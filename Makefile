# Convenience targets; everything also works as the plain commands shown in README.md.
PY ?= python

build:            ## compile the native library for sm_100a (in-tree .so)
	$(PY) -m cpd_b200.build

test:             ## CPU test-suite
	$(PY) -m pytest tests -q -m "not gpu"

test-gpu:         ## GPU test-suite (needs a B200)
	$(PY) -m pytest tests -q -m gpu

sanitize:         ## ASAN + UBSAN + -Werror self-test of the C++ numerics
	bash tools/sanitize_cpu.sh

bench:            ## headline benchmark on one GPU
	$(PY) bench.py --gpus 1

floors:           ## re-check the committed measurements against the performance floors
	$(PY) tools/check_perf_floors.py

docs:             ## regenerate docs/API.md and profiles/sass/RESOURCES.md
	$(PY) tools/gen_api_doc.py
	bash tools/kernel_resources.sh

.PHONY: build test test-gpu sanitize bench floors docs
