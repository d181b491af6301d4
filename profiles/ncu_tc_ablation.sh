cd /root/repo
export CPB200_LIBRARY=/root/repo/channel-pruning_b200/libcpb200_timing.so
for m in 64 95; do
  TC_MASKS=$m ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/tc_mask$m.csv python profiles/tc_timeline.py > /dev/null 2>&1
  echo "mask $m"; grep -E "gram_tc_kernel|reduce_tc|colsum|mirror" gpurun_out/tc_mask$m.csv | awk -F'","' '{print $5, $NF}' | tail -8
done
