#!/usr/bin/env bash
# Shows how this image's nvcc contracts the three source expressions the absent upstream kernels use.
# Run:  bash oracle/probe_fma_contraction.sh      (needs nvcc only, no GPU)
# Result with nvcc 12.9.86 (recorded in DESIGN.md):
#   pointnet2_ops / chamfer  `a*a + b*b + c*c`      -> mul(b,b); fma(a,a,.); fma(c,c,.)
#   KNN_CUDA  `ssd = 0; ssd += t*t` per dim         -> fma(dx,dx,0); fma(dy,dy,.); fma(dz,dz,.)
#   pointnet2_ops  `mag <= 1e-3`                    -> cvt.f64.f32 + setp.le.f64 (double compare)
set -euo pipefail
tmp=$(mktemp -d)
cat > "$tmp/c.cu" <<'CU'
__global__ void fps_expr(const float* p, float* o){
  float x1=p[0],y1=p[1],z1=p[2],x2=p[3],y2=p[4],z2=p[5];
  float d = (x2 - x1) * (x2 - x1) + (y2 - y1) * (y2 - y1) + (z2 - z1) * (z2 - z1);
  float mag = (x2 * x2) + (y2 * y2) + (z2 * z2);
  o[0]=d; o[1]=mag; o[2] = (mag <= 1e-3) ? 1.f : 0.f;
}
__global__ void chamfer_expr(const float* p, float* o){
  float x1=p[0],y1=p[1],z1=p[2];
  float x2=p[3]-x1,y2=p[4]-y1,z2=p[5]-z1;
  o[0]=x2*x2+y2*y2+z2*z2;
}
__global__ void knn_expr(const float* a, const float* b, float* o){
  float ssd=0;
  for(int k=0;k<3;k++){ float tmp=a[k]-b[k]; ssd+=tmp*tmp; }
  o[0]=ssd;
}
CU
nvcc -gencode arch=compute_100a,code=sm_100a -ptx "$tmp/c.cu" -o "$tmp/c.ptx"
grep -E "^\.visible|sub\.f32|mul\.f32|fma\.rn|setp|cvt\.f64" "$tmp/c.ptx"
rm -rf "$tmp"
