#include <thread>
#include <vector>
#include <chrono>
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cstring>
int main(int argc,char**argv){ int T=atoi(argv[1]); size_t n=1ull<<28; uint8_t*buf=(uint8_t*)malloc(n); memset(buf,1,n);
 for(int it=0;it<2;it++){ auto t0=std::chrono::steady_clock::now(); std::vector<std::thread> th; std::vector<uint64_t> sums(T);
 for(int t=0;t<T;t++) th.emplace_back([&,t]{ uint64_t s=0; const uint64_t*p=(const uint64_t*)(buf+n*t/T); size_t m=n/T/8; for(size_t i=0;i<m;i++) s+=p[i]; sums[t]=s;});
 for(auto&x:th)x.join(); double dt=std::chrono::duration<double>(std::chrono::steady_clock::now()-t0).count(); printf("T=%d %.1f GB/s %lu\n",T,n/dt/1e9,(unsigned long)sums[0]);} }
