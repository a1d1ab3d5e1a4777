export GM3D_KNN_XSORT_MIN=1024
for m in 0 1 2; do echo "mode $m"; GM3D_XS_MODE=$m python tools/bench_knn.py 2>&1 | head -1; done
