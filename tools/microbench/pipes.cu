// Microbenchmark: FFMA / FFMA2 / broadcast-LDS issue rates on sm_100a (design input for the TP kernel).
#include <cstdio>
#include <cuda_runtime.h>
#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("ERR %s line %d\n",cudaGetErrorString(e),__LINE__);return 1;}}while(0)

template<int MODE>
__global__ void __launch_bounds__(256) k(float* out, int iters, const float* in) {
  __shared__ __align__(16) float sm[1024];
  for (int i = threadIdx.x; i < 1024; i += blockDim.x) sm[i] = in[i];
  __syncthreads();
  float a[16]; float2 a2[8];
  #pragma unroll
  for (int i=0;i<16;i++) a[i]=in[threadIdx.x%7+i];
  #pragma unroll
  for (int i=0;i<8;i++) a2[i]=make_float2(a[2*i],a[2*i+1]);
  float x0=in[threadIdx.x&31], x1=in[(threadIdx.x&31)+1];
  float2 xx=make_float2(x0,x1);
  int off = (in[3] > 1e30f) ? threadIdx.x : 0;   // runtime-zero, defeats hoisting; warp-uniform address
  for (int it=0; it<iters; ++it) {
    if (MODE==0) { // scalar FFMA, 16 independent chains, reg operands
      #pragma unroll
      for (int r=0;r<4;r++){
        #pragma unroll
        for (int i=0;i<16;i++) a[i]=fmaf(a[i],x0,x1);
      }
    } else if (MODE==1) { // FFMA2
      #pragma unroll
      for (int r=0;r<4;r++){
        #pragma unroll
        for (int i=0;i<8;i++) a2[i]=__ffma2_rn(a2[i],xx,xx);
      }
    } else if (MODE==2) { // 1 LDS.128 broadcast per 4 FFMA
      #pragma unroll
      for (int r=0;r<4;r++){
        #pragma unroll
        for (int q=0;q<4;q++){
          float4 m=*(const float4*)&sm[((it*16+r*4+q)*4+off)&1020];
          a[q*4+0]=fmaf(m.x,x0,a[q*4+0]); a[q*4+1]=fmaf(m.y,x0,a[q*4+1]);
          a[q*4+2]=fmaf(m.z,x0,a[q*4+2]); a[q*4+3]=fmaf(m.w,x0,a[q*4+3]);
        }
      }
    } else if (MODE==3) { // 1 LDS.128 per 8 FFMA (2 channels)
      #pragma unroll
      for (int r=0;r<2;r++){
        #pragma unroll
        for (int q=0;q<4;q++){
          float4 m=*(const float4*)&sm[((it*8+r*4+q)*4+off)&1020];
          a[q*4+0]=fmaf(m.x,x0,a[q*4+0]); a[q*4+1]=fmaf(m.y,x0,a[q*4+1]);
          a[q*4+2]=fmaf(m.z,x0,a[q*4+2]); a[q*4+3]=fmaf(m.w,x0,a[q*4+3]);
          a[q*4+0]=fmaf(m.x,x1,a[q*4+0]); a[q*4+1]=fmaf(m.y,x1,a[q*4+1]);
          a[q*4+2]=fmaf(m.z,x1,a[q*4+2]); a[q*4+3]=fmaf(m.w,x1,a[q*4+3]);
        }
      }
    } else if (MODE==4) { // 1 LDS.128 per 4 FFMA2 (k-paired, 2 channels): 8 FMAs
      #pragma unroll
      for (int r=0;r<4;r++){
        #pragma unroll
        for (int q=0;q<2;q++){
          float4 m=*(const float4*)&sm[((it*8+r*2+q)*4+off)&1020];
          float2 m01=make_float2(m.x,m.y), m23=make_float2(m.z,m.w);
          float2 xa=make_float2(x0,x0), xb=make_float2(x1,x1);
          a2[q*4+0]=__ffma2_rn(m01,xa,a2[q*4+0]); a2[q*4+1]=__ffma2_rn(m23,xa,a2[q*4+1]);
          a2[q*4+2]=__ffma2_rn(m01,xb,a2[q*4+2]); a2[q*4+3]=__ffma2_rn(m23,xb,a2[q*4+3]);
        }
      }
    } else if (MODE==5) { // LDS.128 broadcast only
      #pragma unroll
      for (int r=0;r<16;r++){
        float4 m=*(const float4*)&sm[((it*16+r)*4+off)&1020];
        a[r]+=m.x+m.w; // 2 FADD per LDS
      }
    } else if (MODE==6) { // LDS.32 broadcast : 1 per 1 FFMA
      #pragma unroll
      for (int r=0;r<4;r++){
        #pragma unroll
        for (int i=0;i<16;i++){ float m=sm[((it*64+r*16+i)+off)&1023]; a[i]=fmaf(m,x0,a[i]); }
      }
    } else if (MODE==7) { // FFMA with immediate constant
      #pragma unroll
      for (int r=0;r<4;r++){
        #pragma unroll
        for (int i=0;i<16;i++) a[i]=fmaf(a[i],0.40824829f,x1);
      }
    } else if (MODE==8) { // LDS.64 broadcast per 2 FFMA2 channel-paired w/ duplicated M: m=(M,M)
      #pragma unroll
      for (int r=0;r<4;r++){
        #pragma unroll
        for (int q=0;q<4;q++){
          float4 m=*(const float4*)&sm[((it*16+r*4+q)*4+off)&1020];
          a2[q*2+0]=__ffma2_rn(make_float2(m.x,m.y),xx,a2[q*2+0]);
          a2[q*2+1]=__ffma2_rn(make_float2(m.z,m.w),xx,a2[q*2+1]);
        }
      }
    }
  }
  float s=0;
  #pragma unroll
  for (int i=0;i<16;i++) s+=a[i];
  #pragma unroll
  for (int i=0;i<8;i++) s+=a2[i].x+a2[i].y;
  out[blockIdx.x*blockDim.x+threadIdx.x]=s;
}

template<int MODE> int run(const char* name, double fma_per_iter, float* out, float* in, int nsm) {
  int iters=4096; dim3 g(nsm*4), b(256);
  k<MODE><<<g,b>>>(out,16,in); CK(cudaDeviceSynchronize());
  cudaEvent_t e0,e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  float best=1e30f;
  for (int rep=0;rep<3;rep++){ cudaEventRecord(e0); k<MODE><<<g,b>>>(out,iters,in); cudaEventRecord(e1); CK(cudaEventSynchronize(e1)); float ms; cudaEventElapsedTime(&ms,e0,e1); if(ms<best)best=ms; }
  double fmas = (double)g.x*b.x*iters*fma_per_iter;
  printf("%-44s %8.3f ms  %7.2f TFMA/s  (= %6.1f FMA/clk/SM @1.9GHz)\n", name, best, fmas/best/1e9, fmas/(best*1e-3)/nsm/1.9e9);
  return 0;
}
int main(){
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p,0)); int nsm=p.multiProcessorCount;
  int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  printf("%s SMs=%d clock=%d kHz\n",p.name,nsm,clk);
  float *out,*in; CK(cudaMalloc(&out,nsm*4*256*4)); CK(cudaMalloc(&in,4096*4));
  float h[4096]; for(int i=0;i<4096;i++)h[i]=1e-3f*(i%13); CK(cudaMemcpy(in,h,sizeof(h),cudaMemcpyHostToDevice));
  run<0>("FFMA reg (64/iter)",64,out,in,nsm);
  run<7>("FFMA imm (64/iter)",64,out,in,nsm);
  run<1>("FFMA2 (32 instr = 64 FMA/iter)",64,out,in,nsm);
  run<2>("LDS.128 bcast + 4 FFMA (64 FMA/iter)",64,out,in,nsm);
  run<3>("LDS.128 bcast + 8 FFMA (64 FMA/iter)",64,out,in,nsm);
  run<4>("LDS.128 bcast + 4 FFMA2 (64 FMA/iter)",64,out,in,nsm);
  run<8>("LDS.128 bcast + 2 FFMA2 dupM (64 FMA/iter)",64,out,in,nsm);
  run<6>("LDS.32 bcast + 1 FFMA (64 FMA/iter)",64,out,in,nsm);
  run<5>("LDS.128 bcast only (16 LDS/iter; 'FMA'=LDS)",16,out,in,nsm);
  return 0;
}
