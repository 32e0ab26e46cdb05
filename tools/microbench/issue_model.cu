// Issue-slot model microbenchmark (sm_100a): does FFMA2 leave free issue slots for ALU/LDS work?
// For each "other" instruction type T and count X: loop body = 64 FFMA2 (16 independent pair accumulators x 4)
// + X instructions of type T.  Reports cycles per loop body per warp-scheduler (from clock64) and SM clock (MHz).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ unsigned long long pack2(float a,float b){unsigned long long r; asm("mov.b64 %0,{%1,%2};":"=l"(r):"f"(a),"f"(b)); return r;}
__device__ __forceinline__ void unpack2(unsigned long long v,float&a,float&b){asm("mov.b64 {%0,%1},%2;":"=f"(a),"=f"(b):"l"(v));}

// T: 0 none, 1 FMNMX, 2 IADD3 (add.s32), 3 LDS.128 broadcast, 4 FMNMX3, 5 LDS.32 broadcast, 6 FADD2, 7 scalar FFMA instead of FFMA2 (X ignored)
template<int T,int X>
__global__ void __launch_bounds__(256) k(float* out,const float* in,int iters,long long* cyc,unsigned long long* ns){
  __shared__ float4 sb[256];
  for(int i=threadIdx.x;i<256;i+=blockDim.x) sb[i]=make_float4(in[i&63],in[(i+1)&63],in[(i+2)&63],in[(i+3)&63]);
  __syncthreads();
  unsigned long long acc[16]; unsigned long long a[4]; float b[4];
  float facc[32];
  #pragma unroll
  for(int i=0;i<16;i++){ acc[i]=pack2(in[(threadIdx.x+i)&63],in[(threadIdx.x+i+7)&63]); facc[2*i]=in[(threadIdx.x+i)&63]; facc[2*i+1]=in[(threadIdx.x+i+9)&63]; }
  #pragma unroll
  for(int i=0;i<4;i++){ a[i]=pack2(in[64+i+(threadIdx.x&1)],in[68+i+(threadIdx.x&1)]); b[i]=in[72+i+(threadIdx.x&1)]; }
  float mx[8]; int ia[8];
  #pragma unroll
  for(int i=0;i<8;i++){ mx[i]=in[80+i+(threadIdx.x&1)]; ia[i]=threadIdx.x+i; }
  float4 ld=make_float4(0,0,0,0);
  unsigned long long t0g; asm volatile("mov.u64 %0,%%globaltimer;":"=l"(t0g));
  long long t0=clock64();
  for(int it=0;it<iters;it++){
    #pragma unroll
    for(int u=0;u<4;u++){
      #pragma unroll
      for(int i=0;i<16;i++){
        if(T==7){
          asm volatile("fma.rn.f32 %0,%1,%2,%0;":"+f"(facc[2*i]):"f"(b[u]),"f"(mx[i&7]));
          asm volatile("fma.rn.f32 %0,%1,%2,%0;":"+f"(facc[2*i+1]):"f"(b[u]),"f"(mx[(i+1)&7]));
        } else {
          asm volatile("fma.rn.f32x2 %0,%1,%2,%0;":"+l"(acc[i]):"l"(a[(i+u)&3]),"l"(pack2(b[u],b[u])));
        }
        // interleave X others evenly over the 64 FFMA2
        constexpr int per = X; // total per body
        const int idx=u*16+i;
        if(per>0 && (idx*per)/64 != ((idx+1)*per)/64){
          const int o=(idx*per)/64;
          if(T==1) asm volatile("max.f32 %0,%0,%1;":"+f"(mx[o&7]):"f"(b[o&3]));
          if(T==2) asm volatile("add.s32 %0,%0,%1;":"+r"(ia[o&7]):"r"(ia[(o+1)&7]));
          if(T==3){ float4 v; asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3},[%4];":"=f"(v.x),"=f"(v.y),"=f"(v.z),"=f"(v.w):"r"((unsigned)__cvta_generic_to_shared(&sb[(it+o)&255]))); b[0]=v.x; b[1]=v.y; b[2]=v.z; b[3]=v.w; }
          if(T==4) asm volatile("max.f32 %0,%0,%1,%2;":"+f"(mx[o&7]):"f"(b[o&3]),"f"(b[(o+1)&3]));
          if(T==5){ float v; asm volatile("ld.shared.f32 %0,[%1];":"=f"(v):"r"((unsigned)__cvta_generic_to_shared(&sb[(it+o)&255]))); b[o&3]=v; }
          if(T==6) asm volatile("add.rn.f32x2 %0,%0,%1;":"+l"(acc[o&15]):"l"(a[o&3]));
        }
      }
    }
  }
  long long t1=clock64();
  unsigned long long t1g; asm volatile("mov.u64 %0,%%globaltimer;":"=l"(t1g));
  float s=ld.x+ld.y+ld.z+ld.w;
  #pragma unroll
  for(int i=0;i<16;i++){float x,y; unpack2(acc[i],x,y); s+=x+y+facc[2*i]+facc[2*i+1];}
  #pragma unroll
  for(int i=0;i<8;i++) s+=mx[i]+ia[i];
  out[blockIdx.x*blockDim.x+threadIdx.x]=s;
  if(threadIdx.x==0&&blockIdx.x==0){ cyc[0]=t1-t0; ns[0]=t1g-t0g; }
}
template<int T,int X> void run(const char* name,int sms,float* out,float* in,long long* cyc,unsigned long long* ns){
  const int iters=4000;
  for(int cps=1;cps<=2;cps++){
    k<T,X><<<sms*cps,256>>>(out,in,iters,cyc,ns); cudaDeviceSynchronize();
    cudaEvent_t e0,e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0); k<T,X><<<sms*cps,256>>>(out,in,iters,cyc,ns); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms,e0,e1);
    long long c; unsigned long long n; cudaMemcpy(&c,cyc,8,cudaMemcpyDeviceToHost); cudaMemcpy(&n,ns,8,cudaMemcpyDeviceToHost);
    double warps_per_sched=cps*8/4.0;
    double cyc_per_body=(double)c/iters/warps_per_sched;   // scheduler cycles per loop body of one warp
    double tf=(double)sms*cps*256*iters*64*4/ (ms*1e-3)*1e-12;
    printf("%-14s X=%2d ctas/sm=%d: %.1f sched-cycles/body (64 FFMA2=128 ideal)  clk=%.0f MHz  %.2f TFLOP/s  %.3f ms\n",name,X,cps,cyc_per_body,(double)c/n*1e3,tf,ms);
  }
}
int main(){
  cudaDeviceProp p; cudaGetDeviceProperties(&p,0); int sms=p.multiProcessorCount;
  float *in,*out; long long* cyc; unsigned long long* ns;
  cudaMalloc(&in,4096); cudaMalloc(&out,sms*2*256*4); cudaMalloc(&cyc,8); cudaMalloc(&ns,8);
  float h[1024]; for(int i=0;i<1024;i++) h[i]=1.0f+1e-3f*(i%7); cudaMemcpy(in,h,4096,cudaMemcpyHostToDevice);
  run<0,0>("none",sms,out,in,cyc,ns);
  run<7,0>("scalarFFMAx128",sms,out,in,cyc,ns);
  run<1,8>("FMNMX",sms,out,in,cyc,ns);  run<1,16>("FMNMX",sms,out,in,cyc,ns);  run<1,32>("FMNMX",sms,out,in,cyc,ns); run<1,64>("FMNMX",sms,out,in,cyc,ns);
  run<4,16>("FMNMX3",sms,out,in,cyc,ns); run<4,32>("FMNMX3",sms,out,in,cyc,ns);
  run<2,8>("IADD",sms,out,in,cyc,ns);   run<2,16>("IADD",sms,out,in,cyc,ns);   run<2,32>("IADD",sms,out,in,cyc,ns);  run<2,64>("IADD",sms,out,in,cyc,ns);
  run<3,4>("LDS128",sms,out,in,cyc,ns); run<3,8>("LDS128",sms,out,in,cyc,ns);  run<3,16>("LDS128",sms,out,in,cyc,ns);
  run<5,8>("LDS32",sms,out,in,cyc,ns);  run<5,16>("LDS32",sms,out,in,cyc,ns);
  run<6,4>("FADD2",sms,out,in,cyc,ns);  run<6,8>("FADD2",sms,out,in,cyc,ns);
  return 0;
}
