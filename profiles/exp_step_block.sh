OUT=gpurun_out; mkdir -p $OUT
for v in base sb2 sb4; do
  if [ $v = base ]; then unset S3_LIB_PATH; else export S3_LIB_PATH=$PWD/soap3-dp_b200/libsoap3dp_b200.$v.so; fi
  echo "== $v"; timeout 600 python -m pytest tests/test_dp_gpu.py -m gpu -x -q 2>&1 | tail -2
done
unset S3_LIB_PATH
bash profiles/exp_variants.sh r02h base sb2 sb4 2>&1 | grep -v passed
