// FP32 issue-rate microbenchmark for sm_100a: measures the FFMA / FFMA2 (fma.rn.f32x2) peak that the
// mutual-NN kernel's roofline is normalised against, plus mixes with broadcast LDS.128 and FMNMX3.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp32_peak fp32_peak.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("CUDA error %s at %d\n",cudaGetErrorString(e),__LINE__); return 1;}}while(0)

__device__ __forceinline__ unsigned long long pack2(float a,float b){unsigned long long r; asm("mov.b64 %0,{%1,%2};":"=l"(r):"f"(a),"f"(b)); return r;}
__device__ __forceinline__ void unpack2(unsigned long long v,float&a,float&b){asm("mov.b64 {%0,%1},%2;":"=f"(a),"=f"(b):"l"(v));}
__device__ __forceinline__ unsigned long long ffma2(unsigned long long a,unsigned long long b,unsigned long long c){unsigned long long d; asm("fma.rn.f32x2 %0,%1,%2,%3;":"=l"(d):"l"(a),"l"(b),"l"(c)); return d;}

constexpr int NACC = 16;

// mode 0: scalar FFMA, 2*NACC accumulators.  flops/iter/thread = 2*NACC*2*U
template<int U>
__global__ void __launch_bounds__(256) k_ffma(float* out, const float* in, int iters){
  float acc[2*NACC]; float a[4];
  #pragma unroll
  for(int i=0;i<2*NACC;i++) acc[i]=in[(threadIdx.x+i)&63];
  #pragma unroll
  for(int i=0;i<4;i++) a[i]=in[64+i+(threadIdx.x&1)];
  for(int it=0;it<iters;it++){
    #pragma unroll
    for(int u=0;u<U;u++){
      #pragma unroll
      for(int i=0;i<2*NACC;i++) asm volatile("fma.rn.f32 %0,%1,%2,%0;":"+f"(acc[i]):"f"(a[u&3]),"f"(a[(u+1)&3]));
    }
  }
  float s=0; 
  #pragma unroll
  for(int i=0;i<2*NACC;i++) s+=acc[i];
  out[blockIdx.x*blockDim.x+threadIdx.x]=s;
}

// mode 1: FFMA2 with scalar-broadcast b operand (the .F32 form), NACC pair accumulators
template<int U>
__global__ void __launch_bounds__(256) k_ffma2(float* out, const float* in, int iters){
  unsigned long long acc[NACC]; unsigned long long a[4]; float b[4];
  #pragma unroll
  for(int i=0;i<NACC;i++) acc[i]=pack2(in[(threadIdx.x+i)&63],in[(threadIdx.x+i+7)&63]);
  #pragma unroll
  for(int i=0;i<4;i++){ a[i]=pack2(in[64+i+(threadIdx.x&1)],in[68+i+(threadIdx.x&1)]); b[i]=in[72+i+(threadIdx.x&1)]; }
  for(int it=0;it<iters;it++){
    #pragma unroll
    for(int u=0;u<U;u++){
      #pragma unroll
      for(int i=0;i<NACC;i++) acc[i]=ffma2(a[(i+u)&3],pack2(b[u&3],b[u&3]),acc[i]);
    }
  }
  float s=0;
  #pragma unroll
  for(int i=0;i<NACC;i++){float x,y; unpack2(acc[i],x,y); s+=x+y;}
  out[blockIdx.x*blockDim.x+threadIdx.x]=s;
}

// mode 2: FFMA2 + broadcast LDS.128 every 8 FFMA2 (mutual-NN inner-loop shape) + optional FMNMX3 per 16 FFMA2
template<int WITH_MNMX>
__global__ void __launch_bounds__(256) k_mix(float* out, const float* in, int iters){
  __shared__ float4 sb[512];
  for(int i=threadIdx.x;i<512;i+=blockDim.x) sb[i]=make_float4(in[i&63],in[(i+1)&63],in[(i+2)&63],in[(i+3)&63]);
  __syncthreads();
  unsigned long long acc[NACC]; unsigned long long a[8];
  #pragma unroll
  for(int i=0;i<NACC;i++) acc[i]=pack2(in[(threadIdx.x+i)&63],in[(threadIdx.x+i+7)&63]);
  #pragma unroll
  for(int i=0;i<8;i++) a[i]=pack2(in[64+i+(threadIdx.x&1)],in[72+i+(threadIdx.x&1)]);
  float mx=-1e30f;
  for(int it=0;it<iters;it++){
    #pragma unroll
    for(int j=0;j<8;j++){                       // 8 "columns"
      #pragma unroll
      for(int c=0;c<2;c++){                     // 2 LDS.128 (8 k-values) per column
        float4 b=sb[(it*16+j*2+c)&511];         // warp-uniform address -> broadcast
        acc[2*j]  =ffma2(a[0],pack2(b.x,b.x),acc[2*j]);   acc[2*j+1]=ffma2(a[1],pack2(b.x,b.x),acc[2*j+1]);
        acc[2*j]  =ffma2(a[2],pack2(b.y,b.y),acc[2*j]);   acc[2*j+1]=ffma2(a[3],pack2(b.y,b.y),acc[2*j+1]);
        acc[2*j]  =ffma2(a[4],pack2(b.z,b.z),acc[2*j]);   acc[2*j+1]=ffma2(a[5],pack2(b.z,b.z),acc[2*j+1]);
        acc[2*j]  =ffma2(a[6],pack2(b.w,b.w),acc[2*j]);   acc[2*j+1]=ffma2(a[7],pack2(b.w,b.w),acc[2*j+1]);
      }
      if(WITH_MNMX){ float x,y,z,w; unpack2(acc[2*j],x,y); unpack2(acc[2*j+1],z,w);
        float m; asm volatile("max.f32 %0,%1,%2,%3;":"=f"(m):"f"(x),"f"(y),"f"(z)); asm volatile("max.f32 %0,%1,%2,%0;":"+f"(mx):"f"(m),"f"(w)); }
    }
  }
  float s=mx;
  #pragma unroll
  for(int i=0;i<NACC;i++){float x,y; unpack2(acc[i],x,y); s+=x+y;}
  out[blockIdx.x*blockDim.x+threadIdx.x]=s;
}

template<typename F> float timeit(F f,int reps){
  cudaEvent_t e0,e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  f(); f(); cudaDeviceSynchronize();
  float best=1e30f;
  for(int r=0;r<reps;r++){ cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1); float ms; cudaEventElapsedTime(&ms,e0,e1); if(ms<best)best=ms; }
  return best;
}

int main(){
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p,0));
  int sms=p.multiProcessorCount; int clk=0; cudaDeviceGetAttribute(&clk,cudaDevAttrClockRate,0);
  printf("device %s sms %d clockRate %d kHz\n",p.name,sms,clk);
  float *in,*out; CK(cudaMalloc(&in,4096)); CK(cudaMalloc(&out,sms*8*256*4*4));
  float h[1024]; for(int i=0;i<1024;i++) h[i]=1.0f+1e-3f*(i%7); CK(cudaMemcpy(in,h,4096,cudaMemcpyHostToDevice));
  const int iters=20000;
  for(int cps=1;cps<=4;cps*=2){     // CTAs (256 thr) per SM: 8,16,32 warps/SM
    int grid=sms*cps;
    { float ms=timeit([&]{k_ffma<8><<<grid,256>>>(out,in,iters);},5);
      double fl=(double)grid*256*iters*8*(2*NACC)*2; printf("FFMA   ctas/sm %d: %.3f ms  %.2f TFLOP/s\n",cps,ms,fl/ms*1e-9); }
    { float ms=timeit([&]{k_ffma2<8><<<grid,256>>>(out,in,iters);},5);
      double fl=(double)grid*256*iters*8*NACC*4; printf("FFMA2  ctas/sm %d: %.3f ms  %.2f TFLOP/s\n",cps,ms,fl/ms*1e-9); }
    { float ms=timeit([&]{k_mix<0><<<grid,256>>>(out,in,iters/4);},5);
      double fl=(double)grid*256*(iters/4)*8*16*4; printf("FFMA2+LDS128(1:8) ctas/sm %d: %.3f ms  %.2f TFLOP/s\n",cps,ms,fl/ms*1e-9); }
    { float ms=timeit([&]{k_mix<1><<<grid,256>>>(out,in,iters/4);},5);
      double fl=(double)grid*256*(iters/4)*8*16*4; printf("FFMA2+LDS128+FMNMX3 ctas/sm %d: %.3f ms  %.2f TFLOP/s\n",cps,ms,fl/ms*1e-9); }
  }
  CK(cudaDeviceSynchronize());
  return 0;
}
