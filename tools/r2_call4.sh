#!/bin/bash
# round 2, GPU call 4: full GPU test suite, the default bench line, Hermite / Bloch variants, launch list + one full capture of the bench kernel
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -rs > $O/r2c4_pytest_gpu.log 2>&1; echo "pytest rc $?" | tee -a $O/r2c4_status.txt
timeout 1200 python bench.py > $O/r2c4_bench_n1.json 2> $O/r2c4_bench_n1.err; echo "bench rc $?" | tee -a $O/r2c4_status.txt
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $O/r2c4_bench_ref.json 2> $O/r2c4_bench_ref.err; echo "bench ref rc $?" | tee -a $O/r2c4_status.txt
P="python tools/profile_case.py --state B --tstep 1 --tend 1000 --n 1000000 --reps 2 --jit 1"
{
echo "== TsitPap8 step"; $P --alg TsitPap8
echo "== TsitPap8 hermite"; $P --alg TsitPap8 --output hermite
echo "== TsitPap8 bloch hermite"; $P --alg TsitPap8 --output hermite --chart bloch
Q="python tools/profile_case.py --state A --tstep 0.05 --tend 100 --n 1000000 --reps 2 --jit 1 --tol 1e-12"
echo "== cfg0-like 1e6 traj: TsitPap8 tol 1e-12 ts=0:0.05:100 step"; $Q --alg TsitPap8
echo "== cfg0-like hermite"; $Q --alg TsitPap8 --output hermite
echo "== cfg0-like bloch hermite"; $Q --alg TsitPap8 --output hermite --chart bloch
} > $O/r2c4_variants.log 2>&1
echo "variants done" | tee -a $O/r2c4_status.txt
B="python bench.py --steps 2 --warmup 3 --no-cpu --no-extras --no-e2e"
$B > $O/r2c4_plain_bench.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r2_launches_bench_steps2_warmup3.csv $B > $O/r2c4_ncu_launches.log 2>&1
echo "ncu launches rc $?" | tee -a $O/r2c4_status.txt
C="python bench.py --steps 1 --warmup 1 --no-cpu --no-extras --no-e2e"
$C > $O/r2c4_plain_bench1.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:dtwa_jit_ensemble -s 1 -c 1 -f -o $O/r2_rk4_benchsize $C > $O/r2c4_ncu_rk4.log 2>&1
echo "ncu rk4 benchsize rc $?" | tee -a $O/r2c4_status.txt
ls -la $O/*.ncu-rep | tee -a $O/r2c4_status.txt
