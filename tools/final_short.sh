# Short end-of-round check on one B200: full GPU suite, smoke, C5 bench line (with cpu_baseline), C3 / C4 eager + graph.
mkdir -p gpurun_out/final2; O=gpurun_out/final2
timeout 900 python -m pytest tests -m gpu -x -q > $O/tests.log 2>&1; tail -3 $O/tests.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -1
timeout 400 python bench.py --steps 10 --warmup 3 > $O/bench_c5.json 2> $O/bench_c5.err; python -c "import json; d=json.load(open('$O/bench_c5.json')); print('c5', d['value'], d['ms_per_step'], d['roofline']['achieved'], d['roofline']['frac'], d['e2e']['value'], d['gpu_launches'], d['cpu_baseline']['value'])"
for c in c3 c4; do
  timeout 300 python bench.py --config $c --steps 10 --warmup 3 --no-cpu-baseline > $O/bench_$c.json 2>/dev/null
  timeout 200 python bench.py --config $c --steps 10 --warmup 3 --no-cpu-baseline --graph > $O/bench_${c}_graph.json 2>/dev/null
  python -c "import json; d=json.load(open('$O/bench_$c.json')); g=json.load(open('$O/bench_${c}_graph.json')); print('$c', d['value'], d['ms_per_step'], d['roofline']['achieved'], d['roofline']['frac'], 'graph', g['value'], g['ms_per_step'])"
done
timeout 200 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/c5_step.csv python tools/one_step.py c5 > /dev/null 2>&1
python tools/launch_summary.py $O/c5_step.csv 40 > $O/c5_step.txt; head -8 $O/c5_step.txt
