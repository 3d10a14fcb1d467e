O=gpurun_out/${1:-r2n}; mkdir -p $O
for v in "" _l16 _l32; do
  echo "== pair lanes variant '$v'"
  G2V_LIB_PATH=$PWD/gesture2vec_b200/csrc/libg2v_vq$v.so python tools/timeline.py 400 f32 2>&1 | grep -v -i warn | tail -7
done
python tools/timeline.py 16384 f32 2>&1 | grep -v -i warn | tail -7
python tools/timeline.py 512 bf16 2>&1 | grep -v -i warn | tail -7
timeout 600 python tools/refine_probe.py > $O/refine_probe.log 2>&1; echo "probe rc=$?"; cat $O/refine_probe.log | cut -c1-330
timeout 300 python tools/fold_probe3.py 2>&1 | grep -v -i warn | tail -22
timeout 900 python -m pytest tests -m gpu -q -x > $O/pytest.log 2>&1; echo "pytest rc=$?"; tail -5 $O/pytest.log
