# End-of-round evidence run on one B200 (gpurun): tests, smoke, bench lines of every configuration, sample() timings,
# launch lists and one ncu --set full capture of the tap-loop convolution GEMM.  Outputs under gpurun_out/final/.
mkdir -p gpurun_out/final; O=gpurun_out/final
timeout 900 python -m pytest tests -m gpu -x -q > $O/tests.log 2>&1; tail -3 $O/tests.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 400 python bench.py --steps 10 --warmup 3 --sample > $O/bench_c5.json 2> $O/bench_c5.err; python -c "import json; d=json.load(open('$O/bench_c5.json')); print('c5', d['value'], d['ms_per_step'], d['roofline']['achieved'], d['e2e']['value'], d.get('sample'))"
for c in c1 c2 c3 c4; do
  timeout 300 python bench.py --config $c --steps 10 --warmup 3 > $O/bench_$c.json 2>/dev/null
  timeout 200 python bench.py --config $c --steps 10 --warmup 3 --no-cpu-baseline --graph > $O/bench_${c}_graph.json 2>/dev/null
  python -c "import json; d=json.load(open('$O/bench_$c.json')); g=json.load(open('$O/bench_${c}_graph.json')); print('$c', d['value'], d['ms_per_step'], d['roofline']['achieved'], d['roofline']['frac'], 'graph', g['value'], g['ms_per_step'])"
done
timeout 200 python tools/bench_sample.py c1 c3 c4 2>&1 | grep " n=" > $O/sample.txt; cat $O/sample.txt
for c in c3 c4 c5; do
  timeout 250 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/${c}_step.csv python tools/one_step.py $c > /dev/null 2>&1
  python tools/launch_summary.py $O/${c}_step.csv 40 > $O/${c}_step.txt; head -3 $O/${c}_step.txt
done
timeout 250 ncu --set full --clock-control none --import-source on -k regex:gemm_tc_kernel -c 1 -o $O/conv_gemm python tools/bench_conv.py c4 > /dev/null 2>&1; ls -la $O | tail -5
