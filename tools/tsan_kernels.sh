#!/bin/bash
# Race check of the warp-level kernels WITHOUT a GPU: the kernel sources compiled by g++ with -fsanitize=thread and run with
# one host thread per lane (tests/host_emulation/).  Every cross-lane exchange through shared memory must be ordered by a
# __syncwarp / __syncthreads / shuffle (= a std::barrier here); a phase that relied on lock-step execution would show up as
# a ThreadSanitizer data race.  Usage: tools/tsan_kernels.sh [instances]   (writes profiles/tsan_kernels.txt)
set -e
cd "$(dirname "$0")/.."
B=${1:-2}
mkdir -p /tmp/b200mpc_tsan
for k in ocp_ipm ilqr lmpc sysid; do
  [ -f tests/host_emulation/${k}_host.cpp ] || continue
  g++ -std=c++20 -O1 -g -fsanitize=thread -ffp-contract=off -pthread -shared -fPIC -Wno-unknown-pragmas -I tests/host_emulation \
      tests/host_emulation/${k}_host.cpp -o /tmp/b200mpc_tsan/lib${k}_emu.so
done
TSAN_OPTIONS="halt_on_error=0 report_signal_unsafe=0 history_size=2" LD_PRELOAD=$(g++ -print-file-name=libtsan.so) \
  B200MPC_EMU_LIBDIR=/tmp/b200mpc_tsan python tools/tsan_kernels.py "$B" 2>&1 | tee /tmp/b200mpc_tsan/log.txt | grep -v "^==\|^$" | tail -40
{
  echo "tools/tsan_kernels.sh $B  ($(date -u +%Y-%m-%d))"
  grep -E "^kernel|^summary" /tmp/b200mpc_tsan/log.txt
  echo "ThreadSanitizer data-race reports: $(grep -c 'WARNING: ThreadSanitizer: data race' /tmp/b200mpc_tsan/log.txt || true)"
} > profiles/tsan_kernels.txt
cat profiles/tsan_kernels.txt
