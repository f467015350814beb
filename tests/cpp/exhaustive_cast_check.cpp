// Exhaustive check over ALL 2^32 fp32 bit patterns (tools/exhaustive_cast_check.sh, ~20 s on 16 cores):
//   FormatT<E,M>::q(x) and ::q_wide(x) (the casts the CUDA kernels inline) == quantize_ref(x)  for every pattern, and
//   quantize_ieee(pretruncate<E,M>(x)) == quantize_ref(x) for every finite |x| below the overflow band
// (the second identity is what lets an IEEE hardware convert - F2FP E4M3/E5M2, cvt.rn.f16 / bf16 - emit the reference's
// result after a 6-instruction mask).
#define CPD_NO_CUDA
#include <atomic>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <thread>
#include <vector>
#include "../../cpd_b200/csrc/common/fpformat.h"
static float fb(uint32_t u){ float x; std::memcpy(&x,&u,4); return x; }
static uint32_t bf(float x){ uint32_t u; std::memcpy(&u,&x,4); return u; }
static bool same(float a,float b){ return bf(a)==bf(b) || (std::isnan(a)&&std::isnan(b)); }
template <int E,int M> void run(float limit){
  const cpd::Format fr=cpd::make_format(E,M,cpd::ROUND_REF), fi=cpd::make_format(E,M,cpd::ROUND_IEEE);
  std::atomic<long> bad_q{0}, bad_w{0}, bad_p{0}, n_p{0};
  const int T=std::thread::hardware_concurrency()?std::thread::hardware_concurrency():8;
  std::vector<std::thread> th;
  for(int t=0;t<T;++t) th.emplace_back([&,t]{
    long bq=0,bw=0,bp=0,np=0;
    for(uint64_t u=t; u<(1ull<<32); u+=T){
      const float x=fb((uint32_t)u); const float want=cpd::quantize_ref(x,fr);
      if(!same(cpd::FormatT<E,M>::q(x),want)) ++bq;
      if(!same(cpd::FormatT<E,M>::q_wide(x),want)) ++bw;
      if(std::fabs(x)<limit){ ++np; if(!same(cpd::quantize_ieee(cpd::pretruncate<E,M>(x),fi),want)) ++bp; }
    }
    bad_q+=bq; bad_w+=bw; bad_p+=bp; n_p+=np; });
  for(auto& x:th) x.join();
  std::printf("(%d,%d): all 2^32 patterns: q mismatches %ld, q_wide mismatches %ld, pretruncate+ieee mismatches %ld of %ld below %g\n",E,M,bad_q.load(),bad_w.load(),bad_p.load(),n_p.load(),limit);
}
int main(){ run<4,3>(248.0f); run<5,2>(61440.0f); run<5,10>(65520.0f); run<8,7>(3.3e38f); return 0; }
