#!/bin/bash
# Round-end evidence on one B200 (run on the GPU box): tests, smoke, both bench arms, the ncu launch list of the bench
# command and one ncu --set full capture of a developed-flow step.  Usage: bash profiles/final_capture.sh TAG
TAG=${1:-r2_final}
python -m pytest tests -m gpu -q 2>&1 | tail -3 > gpurun_out/${TAG}_tests.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.log 2>&1
python bench.py --impl reference > gpurun_out/${TAG}_reference.json 2> gpurun_out/${TAG}_reference.err
python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py --steps 2 --warmup 3 --preroll 40 --no-cpu --no-e2e --no-general > /dev/null 2>&1
# one developed-flow step, every kernel with the full set (PRE steps x 13 launches are skipped by count)
PRE=1200 ncu --set full --import-source on --clock-control none --kernel-name-base demangled \
    -k regex:'k_(kappa5|advect5|rhs|jacobi_pk|project4|fct_x5|fct_y5)' -s 9608 -c 16 -o gpurun_out/${TAG}_step -f \
    python profiles/exp_one_kernel.py > gpurun_out/${TAG}_ncu.log 2>&1
tail -2 gpurun_out/${TAG}_tests.log; tail -1 gpurun_out/${TAG}_smoke.log; tail -c 400 gpurun_out/${TAG}_bench.json
