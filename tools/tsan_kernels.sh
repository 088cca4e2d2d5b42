#!/bin/bash
# Race check of the kernels WITHOUT a GPU: the whole C-ABI library compiled by g++ with -fsanitize=thread
# (tests/host_emulation/build_emu_library.py) and run with one host thread per CUDA thread.  Every cross-lane exchange
# through shared memory must be ordered by a __syncwarp / __syncthreads / shuffle (= a std::barrier here); a phase that
# relied on lock-step execution of a warp would show up as a ThreadSanitizer data race.
# Usage: tools/tsan_kernels.sh            (writes profiles/tsan_kernels.txt)
set -e
cd "$(dirname "$0")/.."
LOG=/tmp/b200mpc_tsan_log.txt
TSAN_OPTIONS="halt_on_error=0 report_signal_unsafe=0 history_size=2" LD_PRELOAD=$(g++ -print-file-name=libtsan.so) \
  B200MPC_EMU_TSAN=1 python tools/tsan_kernels.py > $LOG 2>&1 || true
{
  echo "tools/tsan_kernels.sh  ($(date -u +%Y-%m-%d)): libb200mpc_emu.so built with -fsanitize=thread, one host thread per CUDA thread"
  grep -E "^kernel|^summary" $LOG
  echo "ThreadSanitizer data-race reports: $(grep -c 'WARNING: ThreadSanitizer: data race' $LOG || true)"
  grep -A12 'WARNING: ThreadSanitizer: data race' $LOG | grep -E "#[0-3] " | sort | uniq -c | sort -rn | head -20
} > profiles/tsan_kernels.txt
cat profiles/tsan_kernels.txt
