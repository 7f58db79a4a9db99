for per in 20 200 5 20 200; do timeout 200 python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline --clock-period-ms $per > gpurun_out/r02w_clk_$per.json 2>/dev/null; python -c "
import json
d=json.loads(open('gpurun_out/r02w_clk_$per.json').read().strip().splitlines()[-1])
print('period $per', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), 'u8', round(d['e2e_u8']['value'],1), d['clocks'])
"; done 2>&1 | tee gpurun_out/r02w_clock_period.txt
