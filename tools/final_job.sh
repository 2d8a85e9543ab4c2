# the round's measurement call (run on the GPU box through gpurun): tests, both bench arms, launch lists, one full ncu capture
mkdir -p gpurun_out
(time timeout 900 python -m pytest tests -m gpu -x -q) > gpurun_out/r2_tests.log 2>&1
python bench.py --steps 20 --warmup 3 > gpurun_out/bench_r2.json 2> gpurun_out/bench_r2.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_r2_ref.json 2> gpurun_out/bench_r2_ref.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_r2.csv python bench.py --steps 2 --warmup 1 --no-extras --skip-cpu-baseline --sustained-seconds 0 > gpurun_out/r2_ncu_a.log 2>&1
for c in multi8g r9004g; do
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_r2_$c.csv python bench.py --config $c --steps 2 --warmup 1 --no-extras --skip-cpu-baseline --sustained-seconds 0 > gpurun_out/r2_ncu_$c.log 2>&1
done
ncu --set full --clock-control none --import-source on -k regex:demod_fast -s 3 -c 1 -o gpurun_out/prof_r2_final -f python bench.py --steps 2 --warmup 1 --no-extras --skip-cpu-baseline --sustained-seconds 0 > gpurun_out/r2_ncu_full.log 2>&1
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke.log 2>&1
