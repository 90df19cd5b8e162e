#!/bin/bash
# gpurun_out/<tag>_* (written by tools/gpu_round.sh <tag>) -> the tracked summaries under profiles/
TAG=${1:-r1e}; OUT=${2:-r1}
python tools/dram_summary.py gpurun_out/${TAG}_dram_inpipeline.csv > profiles/${OUT}_dram_inpipeline.md
cp gpurun_out/${TAG}_dram_inpipeline.csv profiles/${OUT}_dram_inpipeline.csv
cp gpurun_out/${TAG}_launches.csv profiles/${OUT}_launches.csv
python tools/ncu_summary.py gpurun_out/${TAG}_prof_frame.ncu-rep profiles/ncu_traffic.json > profiles/${OUT}_frame_ncu_full.md
python tools/ncu_summary.py gpurun_out/${TAG}_prof_fullray.ncu-rep profiles/ncu_traffic.json > profiles/${OUT}_fullray_ncu_full.md
tail -1 gpurun_out/${TAG}_bench_fused.json > profiles/${OUT}_bench_fused.json
tail -1 gpurun_out/${TAG}_bench_pingpong.json > profiles/${OUT}_bench_pingpong.json
