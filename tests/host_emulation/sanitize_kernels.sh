#!/bin/bash
# Race and memory check of the kernels WITHOUT a GPU (the host-side counterpart of compute-sanitizer racecheck / memcheck):
# the whole C-ABI library compiled by g++ (tests/host_emulation/build_emu_library.py) with -fsanitize=thread, then with
# -fsanitize=address, and run with one host thread per CUDA thread through the product's batch API.
#  * ThreadSanitizer: every cross-lane exchange through shared memory must be ordered by a __syncwarp / __syncthreads /
#    shuffle (= a std::barrier here); a phase that relied on lock-step execution of a warp shows up as a data race.
#  * AddressSanitizer: "device" buffers are malloc'ed and the dynamic shared memory of each block is allocated with exactly
#    the launch's size, so out-of-bounds global and shared accesses are caught; shared memory starts as NaN bytes, so a read
#    of unwritten shared memory cannot go unnoticed in the results.
# Usage: tests/host_emulation/sanitize_kernels.sh            (writes profiles/sanitize_kernels.txt)
set -e
cd "$(dirname "$0")/../.."
T=/tmp/b200mpc_tsan_log.txt
A=/tmp/b200mpc_asan_log.txt
TSAN_OPTIONS="halt_on_error=0 report_signal_unsafe=0 history_size=2" LD_PRELOAD=$(g++ -print-file-name=libtsan.so) \
  B200MPC_EMU_TSAN=1 python tests/host_emulation/sanitize_kernels.py > $T 2>&1 || true
ASAN_OPTIONS="detect_leaks=0:halt_on_error=0" LD_PRELOAD=$(g++ -print-file-name=libasan.so) \
  B200MPC_EMU_ASAN=1 python tests/host_emulation/sanitize_kernels.py > $A 2>&1 || true
{
  echo "tests/host_emulation/sanitize_kernels.sh  ($(date -u +%Y-%m-%d)): libb200mpc_emu.so = capi.cu + every kernel header compiled by g++, one host thread per CUDA thread"
  echo "--- -fsanitize=thread"
  grep -E "^kernel|^summary" $T
  echo "ThreadSanitizer data-race reports: $(grep -c 'WARNING: ThreadSanitizer: data race' $T || true)"
  grep -A12 'WARNING: ThreadSanitizer: data race' $T | grep -E "#[0-3] " | sort | uniq -c | sort -rn | head -20
  echo "--- -fsanitize=address"
  grep -E "^kernel|^summary" $A
  echo "AddressSanitizer errors: $(grep -c 'ERROR: AddressSanitizer' $A || true)"
  grep -A10 'ERROR: AddressSanitizer' $A | head -30
} > profiles/sanitize_kernels.txt
cat profiles/sanitize_kernels.txt
