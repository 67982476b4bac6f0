#include <pthread.h>
#include <stdio.h>
#include <time.h>
static void* w(void* a){ volatile unsigned long x=0; for(unsigned long i=0;i<1500000000UL;++i) x+=i; return 0;}
int main(int c,char**v){ int n=atoi(v[1]); pthread_t t[64]; struct timespec a,b; clock_gettime(CLOCK_MONOTONIC,&a);
 for(int i=0;i<n;++i) pthread_create(&t[i],0,w,0); for(int i=0;i<n;++i) pthread_join(t[i],0);
 clock_gettime(CLOCK_MONOTONIC,&b); printf("%d threads %.2f s\n", n, (b.tv_sec-a.tv_sec)+(b.tv_nsec-a.tv_nsec)*1e-9); }
