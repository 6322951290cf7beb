# end-of-round verification on one B200: GPU suite, smoke, the default bench line
cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -2 | tr "\n" " "; echo
python __graft_entry__.py smoke 2>&1 | tail -1
timeout 900 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; tail -1 gpurun_out/bench_default.err
python -c "
import json; d=json.load(open('gpurun_out/bench_default.json')); print(round(d['value'],2), 'fps e2e', round(d['e2e']['value'],2), 'frac', round(d['roofline']['frac'],4), 'verified', d['verified'], d['clocks'])"
