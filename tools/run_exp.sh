L=/root/repo/gesture2vec_b200/csrc
echo "== parity"; timeout 40 python tools/tc_debug.py 70000,400,400 300000,512,400 5000,200,448 3000,1000,400 2>&1 | tail -4
for k in 400 512; do timeout 25 python tools/tc_time.py 1000000 $k 400 fast 2>&1 | tail -1 | sed "s/^/fast /"; done
timeout 25 python tools/tc_time.py 1000000 400 400 2>&1 | tail -1 | sed "s/^/full /"
G2V_TC_ABUFS=2 timeout 25 python tools/tc_time.py 1000000 400 400 fast 2>&1 | tail -1 | sed "s/^/abufs2 fast /"
G2V_LIB_PATH=$L/libg2v_vq_t.so timeout 30 python tools/tc_trace.py 1000000 400 400 > gpurun_out/trace_k400_b.txt 2>&1
