// exhaustive check: for every finite float a, is fma(r, y, q0) == a / 25.0f (IEEE), with y = RN(1/25), q0 = RN(a*y), r = fma(-25, q0, a)?
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <pthread.h>
static uint64_t bad[8]; static float first_bad[8]; static float min_ok_abs[8];
static void* run(void* arg){ int id=(int)(intptr_t)arg; const float y = 1.0f/25.0f; uint64_t nb=0; float fb=0; 
  for(uint64_t u=(uint64_t)id; u<0x100000000ull; u+=8){ uint32_t b=(uint32_t)u; float a; memcpy(&a,&b,4); if(!isfinite(a)) continue;
    float q0=a*y; float r=fmaf(-25.0f,q0,a); float q1=fmaf(r,y,q0); float t=a/25.0f; if(memcmp(&q1,&t,4)!=0){ if(fabsf(a) >= 1e-30f){ if(!nb) fb=a; nb++; } } }
  bad[id]=nb; first_bad[id]=fb; return 0; }
int main(){ pthread_t th[8]; for(int i=0;i<8;i++) pthread_create(&th[i],0,run,(void*)(intptr_t)i); uint64_t tot=0; for(int i=0;i<8;i++){ pthread_join(th[i],0); tot+=bad[i]; if(bad[i]) printf("thread %d first bad %g\n", i, first_bad[i]); }
  printf("mismatches with |a| >= 1e-30: %llu\n",(unsigned long long)tot); return 0; }
