// Microbenchmark: DMMA.8x8x4 throughput with the instruction mix of gp_tile_kernel's inner loop.
// Variants add, one at a time: many accumulators, LDS B-fragments, LDG A-fragments, register rotation.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){fprintf(stderr,"CUDA %s at %d\n",cudaGetErrorString(e),__LINE__); exit(1);} }while(0)

__device__ __forceinline__ void dmma(double& c0,double& c1,double a,double b){
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1},{%2},{%3},{%0,%1};":"+d"(c0),"+d"(c1):"d"(a),"d"(b));
}
__device__ __forceinline__ double ldg_stream(const double* p){double v; asm volatile("ld.global.nc.L1::no_allocate.f64 %0, [%1];":"=d"(v):"l"(p)); return v;}

// RQ row blocks x 8 n-blocks per k-step.  MODE bit0: LDS B each step; bit1: LDG A each step (prefetch 3); bit2: rotate regs
template<int RQ,int MODE,int NTH>
__global__ void __launch_bounds__(NTH,1) mix(double* out,const double* W,int iters){
  extern __shared__ double Ks[];
  const int lane=threadIdx.x&31, warp=threadIdx.x>>5;
  for(int i=threadIdx.x;i<256*68;i+=blockDim.x) Ks[i]=1e-3*i;
  __syncthreads();
  double acc[RQ][8][2];
  #pragma unroll
  for(int q=0;q<RQ;q++) for(int nb=0;nb<8;nb++){acc[q][nb][0]=0;acc[q][nb][1]=0;}
  const double* ks_lane=Ks+(lane&3)*68+(lane>>2);
  const double* ap[RQ];
  #pragma unroll
  for(int q=0;q<RQ;q++) ap[q]=W+((size_t)(warp*RQ+q)*4096)*32+lane;
  double a0[RQ],a1[RQ],a2[RQ],b[8];
  #pragma unroll
  for(int q=0;q<RQ;q++){a0[q]=ldg_stream(ap[q]);a1[q]=ldg_stream(ap[q]+32);a2[q]=ldg_stream(ap[q]+64);}
  #pragma unroll
  for(int nb=0;nb<8;nb++) b[nb]=ks_lane[nb*8];
  #pragma unroll 1
  for(int it=0;it<iters;++it){
    const int kk=it&63;
    double a3[RQ];
    if(MODE&2){
      #pragma unroll
      for(int q=0;q<RQ;q++) a3[q]=ldg_stream(ap[q]+((kk+3)&63)*32);
    }
    double bn[8];
    if((MODE&1)&&!(MODE&8)){
      const double* kb=ks_lane+kk*(4*68);
      #pragma unroll
      for(int nb=0;nb<8;nb++) b[nb]=kb[nb*8];
    }
    if(MODE&8){  // B for the NEXT step is loaded while this step's DMMAs run
      const double* kb=ks_lane+((kk+1)&63)*(4*68);
      #pragma unroll
      for(int nb=0;nb<8;nb++) bn[nb]=kb[nb*8];
    }
    #pragma unroll
    for(int q=0;q<RQ;q++){
      #pragma unroll
      for(int nb=0;nb<8;nb++) dmma(acc[q][nb][0],acc[q][nb][1],a0[q],b[nb]);
    }
    if(MODE&8){
      #pragma unroll
      for(int nb=0;nb<8;nb++) b[nb]=bn[nb];
    }
    if(MODE&4){
      #pragma unroll
      for(int q=0;q<RQ;q++){a0[q]=a1[q];a1[q]=a2[q];a2[q]=(MODE&2)?a3[q]:a0[q];}
    }
  }
  double s=0;
  #pragma unroll
  for(int q=0;q<RQ;q++) for(int nb=0;nb<8;nb++) s+=acc[q][nb][0]+acc[q][nb][1];
  if(s==123.456) out[0]=s;
}

// Variant "wide": two k-steps per loop iteration.  A fragments for both k-steps come from one
// 128-bit global load per row block, B fragments from one 128-bit shared load per column block;
// the prefetch ring is unrolled (static registers, no rotation MOVs).  DEPTH = ring depth in
// pairs of k-steps.
__device__ __forceinline__ double2 ldg_stream2(const double2* p){double2 v; asm volatile("ld.global.nc.L1::no_allocate.v2.f64 {%0,%1}, [%2];":"=d"(v.x),"=d"(v.y):"l"(p)); return v;}

template<int RQ,int NTH>
__global__ void __launch_bounds__(NTH,1) mix_wide(double* out,const double* W,int iters){
  extern __shared__ double Ks[];
  const int lane=threadIdx.x&31, warp=threadIdx.x>>5;
  for(int i=threadIdx.x;i<32*4*66*2;i+=blockDim.x) Ks[i]=1e-3*i;
  __syncthreads();
  double acc[RQ][8][2];
  #pragma unroll
  for(int q=0;q<RQ;q++) for(int nb=0;nb<8;nb++){acc[q][nb][0]=0;acc[q][nb][1]=0;}
  // B layout: [pair m][row r=0..3][col 0..65 (66 stride)][2]  -> lane reads double2
  const double2* ks_lane=reinterpret_cast<const double2*>(Ks)+(lane&3)*66+(lane>>2);
  const double2* ap[RQ];
  #pragma unroll
  for(int q=0;q<RQ;q++) ap[q]=reinterpret_cast<const double2*>(W)+((size_t)(warp*RQ+q)*2048)*32+lane;
  double2 ar[3][RQ];
  #pragma unroll
  for(int d=0;d<3;d++)
    #pragma unroll
    for(int q=0;q<RQ;q++) ar[d][q]=ldg_stream2(ap[q]+d*32);
  #pragma unroll 1
  for(int it=0;it<iters;it+=3){
    #pragma unroll
    for(int d=0;d<3;d++){
      const int m=(it+d)&31;
      double2 b[8];
      const double2* kb=ks_lane+m*(4*66);
      #pragma unroll
      for(int nb=0;nb<8;nb++) b[nb]=kb[nb*8];
      #pragma unroll
      for(int q=0;q<RQ;q++){
        #pragma unroll
        for(int nb=0;nb<8;nb++) dmma(acc[q][nb][0],acc[q][nb][1],ar[d][q].x,b[nb].x);
      }
      #pragma unroll
      for(int q=0;q<RQ;q++){
        #pragma unroll
        for(int nb=0;nb<8;nb++) dmma(acc[q][nb][0],acc[q][nb][1],ar[d][q].y,b[nb].y);
      }
      #pragma unroll
      for(int q=0;q<RQ;q++) ar[d][q]=ldg_stream2(ap[q]+((m+3)&31)*32);
    }
  }
  double s=0;
  #pragma unroll
  for(int q=0;q<RQ;q++) for(int nb=0;nb<8;nb++) s+=acc[q][nb][0]+acc[q][nb][1];
  if(s==123.456) out[0]=s;
}

template<int RQ,int WARPS>
double run_wide(int sms,double* out,const double* W){
  const int iters=2049; size_t smem=32*4*66*16;   // iterations count PAIRS of k-steps
  CK(cudaFuncSetAttribute(mix_wide<RQ,WARPS*32>,cudaFuncAttributeMaxDynamicSharedMemorySize,(int)smem));
  cudaEvent_t e0,e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  mix_wide<RQ,WARPS*32><<<sms,WARPS*32,smem>>>(out,W,iters); CK(cudaDeviceSynchronize());
  float best=1e30f;
  for(int r=0;r<5;r++){CK(cudaEventRecord(e0)); mix_wide<RQ,WARPS*32><<<sms,WARPS*32,smem>>>(out,W,iters); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); float ms; CK(cudaEventElapsedTime(&ms,e0,e1)); if(ms<best)best=ms;}
  double fl=2.0*256*RQ*8*2*(double)iters*WARPS*sms;
  return fl/best*1e-9;
}

template<int RQ,int MODE,int WARPS>
double run(int sms,double* out,const double* W){
  const int warps=WARPS;
  const int iters=4096; size_t smem=256*68*8;
  CK(cudaFuncSetAttribute(mix<RQ,MODE,WARPS*32>,cudaFuncAttributeMaxDynamicSharedMemorySize,(int)smem));
  cudaEvent_t e0,e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  mix<RQ,MODE,WARPS*32><<<sms,warps*32,smem>>>(out,W,iters); CK(cudaDeviceSynchronize());
  float best=1e30f;
  for(int r=0;r<5;r++){CK(cudaEventRecord(e0)); mix<RQ,MODE,WARPS*32><<<sms,warps*32,smem>>>(out,W,iters); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); float ms; CK(cudaEventElapsedTime(&ms,e0,e1)); if(ms<best)best=ms;}
  double fl=2.0*256*RQ*8*(double)iters*warps*sms;
  return fl/best*1e-9;
}

int main(){
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p,0)); int sms=p.multiProcessorCount;
  double *out,*W; CK(cudaMalloc(&out,8)); size_t wn=(size_t)16*4*4096*32*2+sms*64*32*2+8192; CK(cudaMalloc(&W,wn*8)); CK(cudaMemset(W,0,wn*8));
  printf("{\"sms\":%d",sms);
  printf(",\"rq4_w8_plain\":%.2f",run<4,0,8>(sms,out,W));
  printf(",\"rq4_w8_lds\":%.2f",run<4,1,8>(sms,out,W));
  printf(",\"rq4_w8_ldg\":%.2f",run<4,2,8>(sms,out,W));
  printf(",\"rq4_w8_rot\":%.2f",run<4,4,8>(sms,out,W));
  printf(",\"rq4_w8_all\":%.2f",run<4,7,8>(sms,out,W));
  printf(",\"rq4_w4_all\":%.2f",run<4,7,4>(sms,out,W));
  printf(",\"rq2_w16_plain\":%.2f",run<2,0,16>(sms,out,W));
  printf(",\"rq2_w16_lds\":%.2f",run<2,1,16>(sms,out,W));
  printf(",\"rq2_w16_all\":%.2f",run<2,7,16>(sms,out,W));
  printf(",\"rq2_w8_all\":%.2f",run<2,7,8>(sms,out,W));
  printf(",\"rq1_w16_all\":%.2f",run<1,7,16>(sms,out,W));
  printf(",\"rq1_w16_plain\":%.2f",run<1,0,16>(sms,out,W));
  printf(",\"rq1_w16_dbuf\":%.2f",run<1,15,16>(sms,out,W));
  printf(",\"rq2_w16_dbuf\":%.2f",run<2,15,16>(sms,out,W));
  printf(",\"rq2_w12_all\":%.2f",run<2,7,12>(sms,out,W));
  printf(",\"rq2_w4_all\":%.2f",run<2,7,4>(sms,out,W));
  printf(",\"rq4_w8_dbuf\":%.2f",run<4,15,8>(sms,out,W));
  printf(",\"rq3_w8_all\":%.2f",run<3,7,8>(sms,out,W));
  printf(",\"rq3_w12_all\":%.2f",run<3,7,12>(sms,out,W));
  printf(",\"wide_rq4_w8\":%.2f",run_wide<4,8>(sms,out,W));
  printf(",\"wide_rq2_w8\":%.2f",run_wide<2,8>(sms,out,W));
  printf(",\"wide_rq1_w8\":%.2f",run_wide<1,8>(sms,out,W));
  printf(",\"wide_rq3_w8\":%.2f",run_wide<3,8>(sms,out,W));
  printf(",\"wide_rq2_w16\":%.2f",run_wide<2,16>(sms,out,W));
  printf("}\n"); return 0;
}
