#include <math.h>
#include <stdio.h>
#include <stdint.h>
#include <string.h>
#include <stdlib.h>
static inline double exp_neg(double x){
  const double MAGIC = 6755399441055744.0;
  double t = fma(x, 1.4426950408889634074, MAGIC);
  int64_t bits; memcpy(&bits,&t,8); int k = (int)(int32_t)(bits & 0xffffffff);
  double kd = t - MAGIC;
  double r = fma(kd, -6.93147180369123816490e-01, x);
  r = fma(kd, -1.90821492927058770002e-10, r);
  double p = 1.0/6227020800.0;         // 1/13!
  p = fma(p, r, 1.0/479001600.0);
  p = fma(p, r, 1.0/39916800.0);
  p = fma(p, r, 1.0/3628800.0);
  p = fma(p, r, 1.0/362880.0);
  p = fma(p, r, 1.0/40320.0);
  p = fma(p, r, 1.0/5040.0);
  p = fma(p, r, 1.0/720.0);
  p = fma(p, r, 1.0/120.0);
  p = fma(p, r, 1.0/24.0);
  p = fma(p, r, 1.0/6.0);
  p = fma(p, r, 0.5);
  p = fma(p, r, 1.0);
  p = fma(p, r, 1.0);
  int64_t pb; memcpy(&pb,&p,8); pb += ((int64_t)k) << 52; memcpy(&p,&pb,8);
  return x < -700.0 ? 0.0 : p;
}
int main(){
  double maxulp=0; int nbad=0; srand(1);
  for(long i=0;i<20000000;i++){
    double u = rand()/(double)RAND_MAX, v=rand()/(double)RAND_MAX;
    double x = -(u*u*u)*50.0 - v*1e-3;   // dense near 0, up to -50
    if(i%7==0) x = -u*700.0;
    double a=exp_neg(x), b=exp(x);
    double ulp = fabs(a-b)/(nextafter(b,INFINITY)-b);
    if(ulp>maxulp) maxulp=ulp; if(ulp>1.0) nbad++;
  }
  printf("max ulp err %.3f, >1ulp: %d; exp_neg(0)=%.17g exp_neg(-1e-300)=%.17g exp_neg(-800)=%g\n",maxulp,nbad,exp_neg(0.0),exp_neg(-1e-300),exp_neg(-800));
  return 0;
}
