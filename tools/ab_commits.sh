set -x
cd ab_old && timeout 300 python bench.py --steps 256 --warmup 8 --no-cpu --no-single > ../gpurun_out/ab_old.json 2> ../gpurun_out/ab_old.err; cd ..
timeout 300 python bench.py --steps 256 --warmup 8 --no-cpu --no-single > gpurun_out/ab_new.json 2> gpurun_out/ab_new.err
cd ab_old && timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu --no-single > ../gpurun_out/ab_old20.json 2>> ../gpurun_out/ab_old.err; cd ..
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu --no-single > gpurun_out/ab_new20.json 2>> gpurun_out/ab_new.err
PROMPT=128 timeout 200 python tools/phase_times.py > gpurun_out/phase128.txt 2>&1
PROMPT=320 timeout 200 python tools/phase_times.py > gpurun_out/phase320.txt 2>&1
cd ab_old; PROMPT=320 timeout 200 python tools/phase_times.py > ../gpurun_out/phase320_old.txt 2>&1; cd ..
python -c "
import json
for f in ['ab_old','ab_new','ab_old20','ab_new20']:
    try:
        d=json.load(open('gpurun_out/'+f+'.json')); print(f, d['value'], d['ms_per_step'], d.get('steady_state'))
    except Exception as e: print(f, 'ERR', e)
"
