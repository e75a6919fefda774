run() { tag=$1; shift; python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $((29600 + RANDOM % 300)) bench.py --gpus 8 --warmup 3 --no-roundtrip --no-cpu --no-fp32 --no-batched "$@" > gpurun_out/s8_$tag.json 2> gpurun_out/s8_$tag.err; python -c "
import json
d=json.loads(open('gpurun_out/s8_$tag.json').read().strip().splitlines()[-1])
print('$tag', round(d['value']/1e9,1), round(d['ms_per_step'],3), round(d['e2e']['value']/1e9,1), d['clocks'].get('samples'))
"; }
run on10 --steps 10
PIC_BENCH_CLOCKS=off run off10 --steps 10
run on10b --steps 10
PIC_BENCH_CLOCKS=off run off10b --steps 10
run on50 --steps 50
