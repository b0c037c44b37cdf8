set -x
O=gpurun_out/r2w; mkdir -p $O
B="python bench.py --streams 1 --steps 2 --warmup 3 --no-cpu-baseline"
ncu --metrics gpu__time_duration.sum --clock-control none -c 1400 --csv --log-file $O/launches.csv $B > $O/bench_under_ncu.json 2> $O/ncu1.err
ncu --set full --clock-control none --import-source on -k regex:k_msm_accumulate --launch-skip 22 -c 7 -o $O/bench_acc $B > /dev/null 2> $O/ncu2.err
ncu --set full --clock-control none -k regex:"k_msm_fold|k_msm_combine|k_msm_final|k_msm_sort" --launch-skip 110 -c 16 -o $O/bench_reduce $B > /dev/null 2> $O/ncu3.err
ncu --set full --clock-control none -k regex:k_ntt_pass --launch-skip 150 -c 12 -o $O/bench_ntt $B > /dev/null 2> $O/ncu4.err
for f in bench_acc bench_reduce bench_ntt; do ncu -i $O/$f.ncu-rep --page raw --csv > $O/${f}_raw.csv 2>/dev/null; done
rm -f $O/bench_reduce.ncu-rep $O/bench_ntt.ncu-rep
# CLI flow with timings
W=$(mktemp -d); mkdir -p $W/data/bfv $W/configs $W/params; cp tests/golden/bfv.in tests/golden/bfv_empty.in $W/data/bfv/
( cd $W; BIN=$GRAFT_REPO_ROOT/zk-fhe_b200/bin/bfv; for c in "setup" "--input bfv/bfv_empty.in keygen" "--input bfv/bfv.in prove" "verify" "--input bfv/bfv.in --transcript blake2b prove" "verify"; do echo "== bfv --name bfv -k 13 $c"; ( time $BIN --name bfv -k 13 $c ); done; echo "== bfv prove --repeat 320 --streams 16 (C++ host threads)"; $BIN --name bfv -k 13 --input bfv/bfv.in --repeat 320 --streams 16 prove; $BIN --name bfv -k 13 --input bfv/bfv.in --transcript blake2b --repeat 320 --streams 16 prove ) > $O/cli.txt 2>&1
# config 5: one limb of the 438-bit RNS modulus at N = 16384, k = 19
python tools/run_config.py --n 16384 --k 19 --rns-bits 438 --limbs 8 --limb 0 --proofs 2 --json $O/config5_limb0_k19.json > $O/config5_limb0_k19.txt 2>&1
# sanitizers on the final code
timeout 600 compute-sanitizer --tool memcheck --log-file $O/memcheck.log python tools/sanitize_run.py --full > $O/memcheck.out 2>&1
timeout 600 compute-sanitizer --tool racecheck --log-file $O/racecheck.log python tools/sanitize_run.py > $O/racecheck.out 2>&1
ZKFHE_NTT_TMA=1 timeout 300 compute-sanitizer --tool memcheck --log-file $O/memcheck_tma.log python tools/sanitize_run.py > $O/memcheck_tma.out 2>&1
# the two arms as the driver runs them
python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > $O/ref_arm.json 2> $O/ref_arm.err
python bench.py --gpus 1 --steps 20 --warmup 5 > $O/bench_arm.json 2> $O/bench_arm.err
tail -c 400 $O/cli.txt; tail -3 $O/config5_limb0_k19.txt; cat $O/memcheck.log $O/racecheck.log $O/memcheck_tma.log | grep -E "ERROR SUMMARY|RACECHECK SUMMARY"
