// write_paths.cpp -- how fast one file takes 4 MB blocks: one writer (mode 0), pwrite at known offsets from T threads (1),
// stores into a shared mapping from T threads (2).  Usage: write_paths FILE MODE T MEGABYTES
#include <fcntl.h>
#include <sys/mman.h>
#include <unistd.h>
#include <string.h>
#include <stdio.h>
#include <stdlib.h>
#include <chrono>
#include <thread>
#include <vector>
#include <atomic>
static double now(){return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();}
int main(int argc,char**argv){
  const char* path=argv[1]; int mode=atoi(argv[2]); int T=atoi(argv[3]); size_t total=(size_t)atof(argv[4])*(1u<<20);
  const size_t BL=4u<<20; size_t nb=total/BL;
  char* src=(char*)malloc(BL); memset(src,'x',BL);
  int fd=open(path,O_RDWR|O_CREAT|O_TRUNC,0666);
  double t0=now();
  if(mode==0){ for(size_t k=0;k<nb;++k) if(pwrite(fd,src,BL,k*BL)!=(ssize_t)BL) return 1; }
  else if(mode==1){ if(ftruncate(fd,total)) return 1; std::atomic<size_t> nx{0}; std::vector<std::thread> th;
    for(int t=0;t<T;++t) th.emplace_back([&]{for(;;){size_t k=nx.fetch_add(1); if(k>=nb)break; if(pwrite(fd,src,BL,k*BL)!=(ssize_t)BL) abort();}});
    for(auto&t:th)t.join(); }
  else if(mode==2){ if(ftruncate(fd,total)) return 1; char* m=(char*)mmap(nullptr,total,PROT_READ|PROT_WRITE,MAP_SHARED,fd,0); if(m==MAP_FAILED){perror("mmap");return 1;}
    std::atomic<size_t> nx{0}; std::vector<std::thread> th;
    for(int t=0;t<T;++t) th.emplace_back([&]{for(;;){size_t k=nx.fetch_add(1); if(k>=nb)break; memcpy(m+k*BL,src,BL);}});
    for(auto&t:th)t.join(); munmap(m,total); }
  double t1=now(); close(fd); double t2=now();
  printf("mode %d T %d: %.3f s (%.2f GB/s), close %.3f\n",mode,T,t1-t0,total/1e9/(t1-t0),t2-t1);
  unlink(path);
}
