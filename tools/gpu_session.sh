#!/bin/bash
# GPU-box session: everything measured in one gpurun call (box acquisition dominates the cost of a call).
#   gpurun --timeout 2400 -- tools/gpu_session.sh <tag> [steps...]      steps default: all
tag=${1:-s}; shift
steps=${@:-pytest smoke bench configs reference clocks ncu counts start launches}
mkdir -p gpurun_out
run() { echo "== $1"; shift; "$@"; echo "   exit $?"; }
for s in $steps; do
case $s in
  pytest)   rm -f gpurun_out/parity_rates.jsonl
            timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -40 > gpurun_out/${tag}_pytest.log; tail -5 gpurun_out/${tag}_pytest.log
            [ -f gpurun_out/parity_rates.jsonl ] && mv gpurun_out/parity_rates.jsonl gpurun_out/${tag}_parity_rates.jsonl ;;
  smoke)    timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${tag}_smoke.log 2>&1; tail -6 gpurun_out/${tag}_smoke.log ;;
  bench)    timeout 400 python bench.py > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; tail -c 600 gpurun_out/${tag}_bench.json ;;
  configs)  for c in 3 4 5; do timeout 400 python bench.py --config $c > gpurun_out/${tag}_bench_c$c.json 2> gpurun_out/${tag}_bench_c$c.err; python - <<PY
import json
try:
    d = json.load(open("gpurun_out/${tag}_bench_c$c.json"))
    print("config $c:", round(d["value"]), "solves/s  e2e", round(d["e2e"]["value"]), " serial", round(d["one_batch_at_a_time"]["value"]), " p50 B=1 %.3f ms" % d["e2e"]["p50_latency_ms_batch1"], " conv", d["converged_frac"], " cpu", d.get("cpu_baseline", {}).get("value"))
except Exception as e:
    print("config $c failed:", e)
PY
            done ;;
  reference) timeout 400 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${tag}_bench_reference.json 2>&1; tail -c 300 gpurun_out/${tag}_bench_reference.json ;;
  clocks)   timeout 300 python tools/phase_clocks.py variants_build/libb200mpc_clocks.so --out gpurun_out/${tag}_phase_clocks.json | cut -c1-400 ;;
  ncu)      timeout 600 ncu --set full --import-source on --clock-control none -k regex:ocp_ipm -s 1 -c 1 -f -o gpurun_out/${tag}_crowded python tools/one_launch.py --B 8192 2>&1 | tail -3
            ls -la gpurun_out/${tag}_crowded.ncu-rep ;;
  counts)   M="smsp__sass_thread_inst_executed_op_dfma_pred_on.sum,smsp__sass_thread_inst_executed_op_dmul_pred_on.sum,smsp__sass_thread_inst_executed_op_dadd_pred_on.sum,smsp__inst_executed.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum"
            timeout 300 ncu --metrics $M --clock-control none -k regex:ocp_ipm -s 1 -c 1 --csv --log-file gpurun_out/${tag}_cnt_cbf.csv python tools/one_launch.py --B 1024 > /dev/null 2>&1
            timeout 300 ncu --metrics $M --clock-control none -k regex:ocp_ipm -s 1 -c 1 --csv --log-file gpurun_out/${tag}_cnt_planner.csv python tools/one_launch.py --B 64 --kind planner > /dev/null 2>&1
            timeout 300 ncu --metrics $M --clock-control none -k regex:lmpc_kernel -s 1 -c 1 --csv --log-file gpurun_out/${tag}_cnt_lmpc.csv python tools/one_launch.py --B 512 --kind lmpc > /dev/null 2>&1
            timeout 300 ncu --metrics $M --clock-control none -k regex:ilqr_kernel -s 1 -c 1 --csv --log-file gpurun_out/${tag}_cnt_ilqr.csv python tools/one_launch.py --B 1024 --kind ilqr > /dev/null 2>&1
            python tools/fp64_counts.py gpurun_out/${tag}_cnt_cbf.csv:1024 gpurun_out/${tag}_cnt_planner.csv:64 gpurun_out/${tag}_cnt_lmpc.csv:512 gpurun_out/${tag}_cnt_ilqr.csv:1024 > gpurun_out/${tag}_fp64_counts.json; head -c 900 gpurun_out/${tag}_fp64_counts.json ;;
  start)    timeout 900 python tools/start_report.py --out gpurun_out/${tag}_start_report.json | tail -1 ;;
  launches) timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches.csv python bench.py --steps 6 --warmup 3 --no-cpu-baseline > /dev/null 2>&1; wc -l gpurun_out/${tag}_launches.csv ;;
  variants) tools/variants.sh run ;;
  sanitize) # compute-sanitizer on small launches of the four solver kernels: memcheck (global / shared out-of-bounds), racecheck
            # (shared-memory hazards between warps / lanes), synccheck (barrier misuse)
            for tool in memcheck racecheck synccheck; do
              for k in "cbf 48" "planner 16" "lmpc 8" "ilqr 32"; do
                set -- $k
                timeout 600 compute-sanitizer --tool $tool --print-limit 5 python tools/one_launch.py --kind $1 --B $2 --reps 1 > gpurun_out/${tag}_sanitizer_${tool}_$1.log 2>&1
                echo "$tool $1 B=$2: $(grep -E "ERROR SUMMARY|RACECHECK SUMMARY" gpurun_out/${tag}_sanitizer_${tool}_$1.log | tail -1)"
              done
            done ;;
  sweep)    timeout 600 python tools/config_sweep.py --no-cpu --out gpurun_out/${tag}_config_sweep.json > /dev/null 2> gpurun_out/${tag}_config_sweep.err; python -c "
import json
d = json.load(open('gpurun_out/${tag}_config_sweep.json'))
print('batch sweep:', [(e['B'], round(e['solves_per_s'])) for e in d['config2_batch_sweep']])
print('overtaking step:', d.get('overtaking_step'))" ;;
  varbench) # every library under scratch/variants swapped in: short bench of configs 2 and 4
            cp car_racing_b200/libb200mpc.so /tmp/libb200mpc_original.so
            for so in variants_build/libb200mpc_*.so; do
              name=$(basename $so .so); name=${name#libb200mpc_}
              cp $so car_racing_b200/libb200mpc.so
              for c in 2 4; do
                timeout 300 python bench.py --config $c --no-cpu-baseline --steps 24 > gpurun_out/${tag}_var_${name}_c$c.json 2>/dev/null
                python -c "
import json
try:
    d = json.load(open('gpurun_out/${tag}_var_${name}_c$c.json')); print('variant $name config $c:', round(d['value']), 'solves/s  serial', round(d['one_batch_at_a_time']['value']), ' p50 B=1 %.3f ms' % d['e2e']['p50_latency_ms_batch1'])
except Exception as e:
    print('variant $name config $c failed', e)"
              done
            done
            cp /tmp/libb200mpc_original.so car_racing_b200/libb200mpc.so ;;
  inflight) for D in 4 8 10; do timeout 300 python bench.py --no-cpu-baseline --inflight $D --steps 48 > gpurun_out/${tag}_bench_D$D.json 2>/dev/null
              python -c "
import json
d = json.load(open('gpurun_out/${tag}_bench_D$D.json')); print('D=$D:', round(d['value']), 'solves/s  e2e', round(d['e2e']['value']))"; done ;;
  latency)  timeout 600 python tools/shim_latency.py --out gpurun_out/${tag}_shim_latency.json > /dev/null 2> gpurun_out/${tag}_shim_latency.err
            python -c "
import json
d = json.load(open('gpurun_out/${tag}_shim_latency.json'))
print('mpccbf', d['mpccbf_test']['latency'], d['mpccbf_test']['statuses'], d['mpccbf_test']['ego_passed_rivals'])
print('mpc_lti', d['mpc_lti_tracking']['latency']['p50_ms'], 'ilqr', d['ilqr_test']['latency']['p50_ms'], 'lmpc step', d['lmpc_step']['both']['p50_ms'], 'overtake', d['overtaking_step']['fused']['p50_ms'], d['overtaking_step']['two_calls']['p50_ms'])" ;;
  ncusum)   timeout 600 ncu --set full --import-source on --clock-control none -k regex:ocp_ipm -s 1 -c 1 -f -o gpurun_out/${tag}_crowded python tools/one_launch.py --B 8192 2>&1 | tail -1
            timeout 600 ncu --set full --import-source on --clock-control none -k regex:lmpc_kernel -s 1 -c 1 -f -o gpurun_out/${tag}_lmpc4096 python tools/one_launch.py --B 4096 --kind lmpc 2>&1 | tail -1
            ls -la gpurun_out/${tag}_*.ncu-rep ;;
  multi)    # N GPUs of this box (gpurun --gpus N): the contract's launch line, configs 2, 4, 5; both forms of the exchange for config 2
            NG=$(nvidia-smi -L | wc -l)
            for c in 2 4 5; do
              timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $NG --config $c --no-cpu-baseline > gpurun_out/${tag}_bench_${NG}gpu_c$c.json 2> gpurun_out/${tag}_bench_${NG}gpu_c$c.err
              python - <<PY
import json
try:
    d = json.load(open("gpurun_out/${tag}_bench_${NG}gpu_c$c.json"))
    print("N=$NG config $c:", round(d["value"]), "solves/s  ms/step %.3f" % d["ms_per_step"], " e2e", round(d["e2e"]["value"]), " exchange:", d["config"]["exchange"][:40], " verified:", d.get("exchange_verified"))
except Exception as e:
    print("N=$NG config $c failed:", e)
PY
            done
            timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $NG --exchange nccl --no-cpu-baseline > gpurun_out/${tag}_bench_${NG}gpu_c2_nccl.json 2> gpurun_out/${tag}_bench_${NG}gpu_c2_nccl.err
            python -c "
import json
d = json.load(open('gpurun_out/${tag}_bench_${NG}gpu_c2_nccl.json')); print('N=$NG config 2 nccl exchange:', round(d['value']), 'solves/s  e2e', round(d['e2e']['value']))" ;;
  multiD)   # config 2 on all GPUs of the box with more batches in flight (slack for the slowest shard)
            NG=$(nvidia-smi -L | wc -l)
            for D in 8 10; do
              timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $NG --inflight $D --no-cpu-baseline > gpurun_out/${tag}_bench_${NG}gpu_c2_D$D.json 2> gpurun_out/${tag}_bench_${NG}gpu_c2_D$D.err
              python -c "
import json
d = json.load(open('gpurun_out/${tag}_bench_${NG}gpu_c2_D$D.json')); print('N=$NG config 2 D=$D:', round(d['value']), 'solves/s  ms/step %.3f' % d['ms_per_step'], ' e2e', round(d['e2e']['value']), d.get('exchange_verified'))"
            done ;;
  sameshards) NG=$(nvidia-smi -L | wc -l)
            timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus $NG --same-shards --no-cpu-baseline > gpurun_out/${tag}_bench_${NG}gpu_c2_sameshards.json 2> gpurun_out/${tag}_bench_${NG}gpu_c2_sameshards.err
            timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29515 bench.py --gpus $NG --no-cpu-baseline > gpurun_out/${tag}_bench_${NG}gpu_c2.json 2> gpurun_out/${tag}_bench_${NG}gpu_c2.err
            timeout 150 python bench.py --no-cpu-baseline > gpurun_out/${tag}_bench_1gpu_c2.json 2>/dev/null
            python -c "
import json
a = json.load(open('gpurun_out/${tag}_bench_${NG}gpu_c2_sameshards.json')); b = json.load(open('gpurun_out/${tag}_bench_${NG}gpu_c2.json')); c = json.load(open('gpurun_out/${tag}_bench_1gpu_c2.json'))
print('N=$NG same shards:', round(a['value']), ' own shards:', round(b['value']), ' 1 GPU same box:', round(c['value']), ' efficiency %.3f / %.3f' % (a['value'] / ($NG * c['value']), b['value'] / ($NG * c['value'])))" ;;
  single)   timeout 400 python bench.py --no-cpu-baseline > gpurun_out/${tag}_bench_1gpu.json 2> gpurun_out/${tag}_bench_1gpu.err
            python -c "
import json
d = json.load(open('gpurun_out/${tag}_bench_1gpu.json')); print('N=1 config 2:', round(d['value']), 'solves/s  e2e', round(d['e2e']['value']))" ;;
  *) echo "unknown step $s" ;;
esac
done
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,temperature.gpu,power.draw --format=csv,noheader
