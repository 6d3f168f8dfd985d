__global__ void k(unsigned long long* a, double* d){
  unsigned long long x=a[threadIdx.x], y=a[threadIdx.x+32];
  double p=d[threadIdx.x], q=d[threadIdx.x+32];
  double hi=__fma_rz(p,q,0x1p104); double lo=__fma_rz(p,q,(0x1p104+0x1p52)-hi);
  x+=__double_as_longlong(hi); y+=__double_as_longlong(lo);
  a[threadIdx.x]=x+y*3;
}
