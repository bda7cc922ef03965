#!/bin/bash
# diagnostics: why does ncu segfault on the GPU box?
set +e
exec > gpurun_out/ncu_diag.log 2>&1
echo "== env"; which ncu python; ncu --version | tail -1; echo HOME=$HOME TMPDIR=$TMPDIR; ls -ld /tmp /tmp/nsight* 2>&1; df -h /tmp | tail -1; ulimit -a | head -20
echo "== 1 torch tiny"
ncu --metrics gpu__time_duration.sum -c 2 python -c "import torch; x=torch.zeros(10,device='cuda'); x+=1; torch.cuda.synchronize(); print('ok')"; echo rc=$?
echo "== 2 torch tiny clock-control none"
ncu --metrics gpu__time_duration.sum --clock-control none -c 2 python -c "import torch; x=torch.zeros(10,device='cuda'); x+=1; torch.cuda.synchronize(); print('ok')"; echo rc=$?
echo "== 3 profile_step 100 reads"
ncu --metrics gpu__time_duration.sum --clock-control none -c 5 python tools/profile_step.py 100 1; echo rc=$?
echo "== 4 csv log-file"
ncu --metrics gpu__time_duration.sum --clock-control none -c 5 --csv --log-file gpurun_out/diag4.csv python tools/profile_step.py 100 1; echo rc=$?
echo "== 5 absolute python"
ncu --metrics gpu__time_duration.sum --clock-control none -c 5 $(readlink -f $(which python)) tools/profile_step.py 100 1; echo rc=$?
echo "== dmesg"; dmesg 2>/dev/null | tail -5
