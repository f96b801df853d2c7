set -u
python -m pytest tests -x -q -m gpu > gpurun_out/t90_tests.log 2>&1
rc=$?
tail -3 gpurun_out/t90_tests.log
if [ $rc -ne 0 ]; then echo "TESTS FAILED"; exit 1; fi
python bench.py --workload illumina_qual_o1 --no-cpu > gpurun_out/b90_o1.json 2> gpurun_out/b90_o1.err; python scripts/benchsum.py gpurun_out/b90_o1.json | head -2
python bench.py --workload illumina_qual_o0 --no-cpu --steps 2 --warmup 1 > gpurun_out/p90_plain.json 2> gpurun_out/p90_plain.err && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/p90_launches_o0.csv python bench.py --workload illumina_qual_o0 --no-cpu --steps 2 --warmup 1 > gpurun_out/p90_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"enc_kernel|dec_kernel|hist_kernel" -s 4 -c 4 -o gpurun_out/p90_o0 -f python bench.py --workload illumina_qual_o0 --no-cpu --steps 1 --warmup 1 > gpurun_out/p90_full.log 2>&1
for w in illumina_qual_o0 illumina_qual_o1; do
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/p90_traffic_$w.csv python scripts/traffic_probe.py $w > gpurun_out/p90_traffic_$w.log 2>&1
done
timeout 600 python bench.py > gpurun_out/b90_all.json 2> gpurun_out/b90_all.err
echo "bench rc=$?"
python scripts/benchsum.py gpurun_out/b90_all.json | head -30
