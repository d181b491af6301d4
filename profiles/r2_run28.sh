#!/bin/bash
# Round-2 GPU call 28: complete GPU test suite (incl. full-size parity) at HEAD, NUMA probe, config-5 sweep, config-4 workload
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
echo "== numa"; python -c "
import cpb200, torch
print('gpu0 numa cpus:', None if cpb200.engine.gpu_numa_cpus(0) is None else len(cpb200.engine.gpu_numa_cpus(0)))
pr = torch.cuda.get_device_properties(0); print(pr.pci_domain_id, pr.pci_bus_id, pr.pci_device_id)" 2>&1 | tail -2 | tee gpurun_out/r3b_numa.log
echo "== tests (all)"; timeout 1800 python -m pytest tests -m gpu -q -x 2>&1 | tail -4 | tee gpurun_out/r3b_tests.log
echo "== sweep"; timeout 600 python bench.py --workload sweep 2>gpurun_out/r3b_sweep.err | tail -1 > gpurun_out/r3b_sweep.json; python -c "
import json; d=json.load(open('gpurun_out/r3b_sweep.json'))
for p in d['points']: print({k:(round(v,4) if isinstance(v,float) else v) for k,v in p.items() if k in ('N','gram_ms','gram_tflops','gram_frac_of_bf16_peak','gather_ms','select_ms','ls_ms')})" 2>&1 | tail -6
echo "== resnet50"; timeout 600 python bench.py --workload resnet50 --steps 3 --warmup 2 --no-cpu --no-e2e 2>&1 | tail -1 | tee gpurun_out/r3b_resnet50.json | cut -c1-260
