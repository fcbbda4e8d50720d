// Host-side microbenchmark of the result-container fill (5.8 M terms into a node pool + bucket heads): what bounds
// hash_set::_bulk_unique_fill.  mode 0 = atomic exchange on random buckets, 1 = input sorted by bucket, 3 = no atomics.
// g++ -std=c++17 -O2 -pthread -o /tmp/fill_bench tools/fill_bench.cpp && /tmp/fill_bench <threads> <mode>
#include <chrono>
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <random>
#include <thread>
#include <vector>
#include <algorithm>
#include <sys/mman.h>
struct node { uint64_t m[2]; uint32_t size; bool neg; int64_t key; uint32_t next; bool live; };
static void *big(size_t b){ void*p=nullptr; size_t r=(b+(2u<<20)-1)&~size_t((2u<<20)-1); posix_memalign(&p,2u<<20,r); madvise(p,r,MADV_HUGEPAGE); return p; }
int main(int argc,char**argv){
  const size_t n=5821335; unsigned nt=argc>1?atoi(argv[1]):8; int mode=argc>2?atoi(argv[2]):0;
  std::vector<int64_t> key(n); std::vector<uint64_t> mag(2*n); std::vector<int8_t> sign(n,1);
  std::mt19937_64 rng(1); int64_t k=-400000000000ll;
  for(size_t i=0;i<n;++i){k+=1+(rng()%200000); key[i]=k; mag[2*i]=rng(); mag[2*i+1]=rng()&0xFFFFFF;}
  size_t nb=1; while(nb<n) nb<<=1;
  if(mode==1){ // bucket-sorted input
    std::vector<size_t> idx(n); for(size_t i=0;i<n;++i) idx[i]=i;
    std::sort(idx.begin(),idx.end(),[&](size_t a,size_t b){return (key[a]&(nb-1))<(key[b]&(nb-1));});
    std::vector<int64_t> k2(n); std::vector<uint64_t> m2(2*n);
    for(size_t i=0;i<n;++i){k2[i]=key[idx[i]]; m2[2*i]=mag[2*idx[i]]; m2[2*i+1]=mag[2*idx[i]+1];}
    key.swap(k2); mag.swap(m2);
  }
  printf("sizeof(node)=%zu\n",sizeof(node));
  for(int rep=0;rep<4;++rep){
    auto t0=std::chrono::steady_clock::now();
    node*pool=(node*)big(n*sizeof(node)); uint32_t*heads=(uint32_t*)big(nb*4);
    auto t1=std::chrono::steady_clock::now();
    auto run=[&](auto&&f){std::vector<std::thread> th; for(unsigned t=1;t<nt;++t) th.emplace_back(f,t); f(0u); for(auto&x:th)x.join();};
    run([&](unsigned t){size_t lo=nb*t/nt,hi=nb*(t+1)/nt; std::fill(heads+lo,heads+hi,0xFFFFFFFFu);});
    auto t2=std::chrono::steady_clock::now();
    run([&](unsigned t){size_t lo=n*t/nt,hi=n*(t+1)/nt;
      for(size_t i=lo;i<hi;++i){
        if(mode!=2 && i+16<hi) __builtin_prefetch(heads+((size_t)key[i+16]&(nb-1)),1,1);
        node&nd=pool[i]; nd.m[0]=mag[2*i]; nd.m[1]=mag[2*i+1]; nd.size=2; nd.neg=false; nd.key=key[i]; nd.live=true;
        size_t b=(size_t)key[i]&(nb-1);
        if(mode==3){ nd.next=heads[b]; heads[b]=(uint32_t)i; }
        else nd.next=__atomic_exchange_n(heads+b,(uint32_t)i,__ATOMIC_RELAXED);
      }});
    auto t3=std::chrono::steady_clock::now();
    free(pool); free(heads);
    auto t4=std::chrono::steady_clock::now();
    auto ms=[](auto a,auto b){return std::chrono::duration<double,std::milli>(b-a).count();};
    printf("mode %d threads %u: alloc %.2f heads %.2f build %.2f free %.2f total %.2f ms\n",mode,nt,ms(t0,t1),ms(t1,t2),ms(t2,t3),ms(t3,t4),ms(t0,t4));
  }
}
