for u in 2 4 14; do
  for a in 2 8 32; do
    [ $u != 2 ] && [ $a != 8 ] && continue
    touch bioinformatics-algorithms_b200/csrc/b2a_api.cu
    make -C bioinformatics-algorithms_b200 all EXTRA_NVFLAGS="-DWIDE_MID_UNROLL=$u -DAFFINE_UNROLL=$a" > /dev/null 2>&1 || echo BUILD FAILED
    echo "== WIDE_MID_UNROLL=$u AFFINE_UNROLL=$a"
    timeout 700 python scripts/bench_long.py --len 100000 --check 0 2>&1 | grep "^config" | grep -v tandem | cut -c1-150
  done
done
