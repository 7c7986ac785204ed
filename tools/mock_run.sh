#!/bin/bash
# Run GPU-marked tests against the CPU mock device (tests/cpp/mock: the host runtime built against a stand-in cuda_runtime.h, kernels replaced by the
# device node library compiled for the CPU) — the same build tests/test_mock_bank_cpu.py makes, for one selection of tests.
# usage: tools/mock_run.sh tests/test_gpu_wider.py -k slot
set -e
ROOT=$(cd "$(dirname "$0")/.." && pwd); cd "$ROOT"
B=tests/cpp/mock/_build/manual; mkdir -p $B
CSRC=fundsp_b200/csrc
g++ -std=c++17 -O1 -ffp-contract=off -w -shared -fPIC -DFDSP_HOST_EMUL=1 -I tests/cpp/mock -I $CSRC -x c++ $CSRC/host/graph.cpp $CSRC/host/wavetable.cpp $CSRC/host/bank.cpp \
    $CSRC/host/group.cpp $CSRC/host/wavfile.cpp $CSRC/capi.cpp tests/cpp/mock/registry_mock.cpp -o $B/libfundsp_b200_mock.so -ldl
FDSP_B200_LIB=$ROOT/$B/libfundsp_b200_mock.so FDSP_MOCK_ROOT=$ROOT FDSP_MOCK_CACHE=$ROOT/$B/classes python -m pytest "$@" -m gpu -q -p no:cacheprovider --tb=short
