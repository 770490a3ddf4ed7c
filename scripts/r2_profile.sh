#!/bin/bash
# round 2: one ncu --set full capture per dominant kernel at a known unit count; summaries -> profiles/r02_*.txt, counters -> profiles/ncu_traffic.json
TAG=${1:-r02}
mkdir -p gpurun_out profiles
cap() { # part kernel-regex units summary-name
  part=$1; rx=$2; n=$3; name=$4
  timeout 900 ncu --set full --metrics l1tex__t_bytes.sum,lts__t_bytes.sum,sm__inst_executed_pipe_fp64.sum --clock-control none --import-source on -k regex:$rx -c 1 -f -o gpurun_out/${TAG}_$name python scripts/prof_part.py $part $n 2 > gpurun_out/${TAG}_ncu_$name.log 2>&1
  python scripts/ncu_summary.py rep gpurun_out/${TAG}_$name.ncu-rep gpurun_out/${TAG}_$name.txt
  python scripts/ncu_summary.py json gpurun_out/${TAG}_$name.ncu-rep $5 $n profiles/${TAG}_$name.txt gpurun_out/${TAG}_ncu_traffic.json
  grep -E "gpu__time_duration|lanes|thread_inst_executed_per|fp64_cycles|l1tex__throughput|dram__bytes" gpurun_out/${TAG}_$name.txt | head -8
}
cap envelope env_points_kernel 10000000 env_points_kernel envelope
cap faces env_faces_kernel 100000 env_faces_kernel envelope_faces
cap nearest nearest_packet_kernel 10000000 nearest_packet_kernel nearest
cap amips amips_soa_tma_kernel 16000000 amips_soa_tma_kernel amips
cap quality mesh_quality_kernel 15000000 mesh_quality_kernel amips_quality
cap ring amips_ring_kernel 16000000 amips_ring_kernel amips_ring
cap winding winding_kernel 2000000 winding_kernel winding
cat gpurun_out/${TAG}_ncu_traffic.json | head -60
