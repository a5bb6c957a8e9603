#!/bin/bash
# round 2 evidence pass on one B200: GPU tests, both bench arms, launch list, ncu captures of every kernel of a step, cycle counters
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
python -m pytest tests -m gpu -q > $O/r2_final_gpu_tests.log 2>&1; tail -3 $O/r2_final_gpu_tests.log
python bench.py > $O/r2_final_bench.json 2> $O/r2_final_bench.err; tail -c 300 $O/r2_final_bench.json; echo
python bench.py --impl reference --steps 1 --warmup 0 > $O/r2_final_bench_reference.json 2> $O/r2_final_bench_reference.err; tail -c 200 $O/r2_final_bench_reference.json; echo
NERFATTN_NO_GRAPH=1 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 120 --csv \
  --log-file $O/r2_final_launches.csv python bench.py --steps 1 --warmup 0 --epochs 3 --no-e2e --no-extras > $O/r2_final_launches.log 2>&1
for a in "chain256 chain_kernel medium 120" "chain512 chain_kernel large 40" "chain256deep chain_kernel deep 40" "dwadam256 dw_adam_kernel medium 120" "dwadam512 dw_adam_kernel large 40" "resident128 resident_kernel small 40 0" "resident64 resident_kernel tiny 40 0"; do
  set -- $a
  bash profiles/cap.sh r2_final_$1 $2 $3 $4 ${5:-2}
done
rm -f $O/r2_final_dwadam*.ncu-rep $O/r2_final_chain256deep.ncu-rep $O/r2_final_resident64.ncu-rep
for a in "medium 120" "large 40" "deep 40"; do set -- $a; bash profiles/timing.sh $1 $2 > $O/r2_final_timing_$1.log 2>&1; tail -3 $O/r2_final_timing_$1.log; done
ls -la $O/r2_final_*
