#!/bin/bash
# A/B harness for kernel experiments under a tight GPU budget: variants are compiled HERE (nvcc cross-compiles without a
# GPU, free) and travel with the gpurun snapshot; ONE gpurun call then times all of them back to back.
#
#   tools/variants.sh build  name1:"-DFOO"  name2:"-DBAR=2 -DBAZ"     # here: variants_build/libb200mpc_<name>.so (+ "base")
#   gpurun -- tools/variants.sh run                                    # on the box: per variant parity subset + timings
#
# `run` swaps each variant into car_racing_b200/libb200mpc.so (the product loads only that path), runs
#   - the config-2 parity tests (-m gpu -k "config2 or shapes or anchor or blocked"; VARIANT_TESTS overrides the selection),
#   - a crowded launch (B=8192), a lone instance (B=1), and a short bench without the CPU leg,
# prints one table row per variant, writes gpurun_out/variants.txt and restores the original library.
# Before building a variant, check it on the host first: B200MPC_EMU_CXXFLAGS="<same -D flags>" python -m pytest
# tests/test_library_on_host.py   (logic on the host-compiled library), and compare `cuobjdump -sass` instruction counts.
set -e
cd "$(dirname "$0")/.."
VDIR=variants_build   # git-ignored, NOT gpurun-ignored: the variant libraries travel with the snapshot
LIB=car_racing_b200/libb200mpc.so
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
FLAGS="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -shared -Xcompiler -fPIC"
case "$1" in
  build)
    shift
    mkdir -p $VDIR
    build_one() {   # name, defs: all translation units in parallel (__graft_entry__.compile_lib)
      python -c "
import sys, __graft_entry__ as g
g.compile_lib(out='$VDIR/libb200mpc_$1.so', extra_flags=sys.argv[1].split(), verbose=False)" "$2"
      n=$(cuobjdump -sass -fun '_ZN7b200mpc14ocp_ipm_kernelILi3ELi8ELi20EEEvNS_7KParamsEPKdP14b200mpc_recordPdS6_S6_S6_NS_8XchgArgsE' $VDIR/libb200mpc_$1.so 2>/dev/null | grep -cE "^\s+/\*[0-9a-f]{4,5}\*/" || true)
      echo "[variants] $1: ocp_ipm_kernel<3,QDIAG,20> = $n SASS instructions"
    }
    [ "$NO_BASE" = "1" ] || build_one base ""
    for spec in "$@"; do
      name=${spec%%:*}; defs=${spec#*:}
      echo "[variants] $name: $defs"
      build_one "$name" "$defs"
    done
    ;;
  run)
    mkdir -p gpurun_out
    cp $LIB /tmp/libb200mpc_original.so
    trap 'cp /tmp/libb200mpc_original.so '$LIB EXIT
    : > gpurun_out/variants.txt
    for so in $VDIR/libb200mpc_*.so; do
      name=$(basename $so .so); name=${name#libb200mpc_}
      cp $so $LIB
      par=$(timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "${VARIANT_TESTS:-config2 or shapes or anchor or blocked}" 2>&1 | tail -1)
      crowded=$(timeout 120 python tools/one_launch.py --B 8192 --reps 3 2>&1 | tail -1)
      lone=$(timeout 60 python tools/one_launch.py --B 1 --reps 3 2>&1 | tail -1)
      timeout 200 python bench.py --no-cpu-baseline --steps 24 > gpurun_out/variant_${name}_bench.json 2>/dev/null || true
      bench=$(python -c "
import json
try:
    d = json.load(open('gpurun_out/variant_${name}_bench.json'))
    print('bench %d solves/s, e2e %d, serial %d, p50 B=1 %.3f ms' % (d['value'], d['e2e']['value'], d['one_batch_at_a_time']['value'], d['e2e']['p50_latency_ms_batch1']))
except Exception as e:
    print('bench failed:', e)")
      { echo "== $name"; echo "   parity : $par"; echo "   crowded: $crowded"; echo "   lone   : $lone"; echo "   $bench"; } | tee -a gpurun_out/variants.txt
    done
    ;;
  *)
    echo "usage: $0 build name:\"-Dflags\" ... | run"; exit 2;;
esac
