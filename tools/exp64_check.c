/* Developer check of the table exp2 used by the Gram matrix and the cross-covariances (bode_b200/csrc/rbf.h, rbf_exp64):
 * the same arithmetic restated in C (fma from libm), compared with exp2l over 2e7 random arguments in [-60000, 0].
 * gcc -O2 -o exp64_check exp64_check.c -lm && ./exp64_check   ->   max rel err 2.273e-16 (1.02 ulp of 2^-52) */
#include <math.h>
#include <stdio.h>
#include <stdint.h>
#include <string.h>
#include <stdlib.h>
static double tab[64];
static double rbf_exp64(double z){
  const double SHIFT = 6755399441055744.0;
  const double t = z + SHIFT; const double kf = t - SHIFT; const double f = z - kf;
  uint64_t tb; memcpy(&tb,&t,8); int k=(int)(uint32_t)tb;
  double p = fma(f, 1.2417843701716925e-12, 5.732851688640402e-10);
  p = fma(p, f, 2.1173137155464776e-07); p = fma(p, f, 5.86490495505617e-05); p = fma(p, f, 0.010830424696249145);
  const double q = p*f; const double T = tab[k&63]; double res=fma(T,q,T);
  int e = k>>6; if(e<-1000)e=-1000;
  uint64_t rb; memcpy(&rb,&res,8); uint32_t hi=(uint32_t)(rb>>32)+((uint32_t)e<<20); rb=((uint64_t)hi<<32)|(rb&0xffffffffu); memcpy(&res,&rb,8); return res;
}
int main(){ for(int j=0;j<64;j++) tab[j]=(double)exp2l((long double)j/64.0L);
  double maxerr=0; srand(1);
  for(long i=0;i<20000000;i++){ double z=-((double)rand()/RAND_MAX)*60000.0*((i&1)?1.0:0.001); long double ref=exp2l((long double)z/64.0L); double r=rbf_exp64(z); double err=fabs((double)((r-ref)/ref)); if(err>maxerr)maxerr=err; }
  printf("max rel err %.3e (ulp=%.3e)\n",maxerr,ldexp(1,-53)); 
  printf("z=1e-12 -> %.17g ; z=0 -> %.17g; z=-0.5 %.17g vs %.17g\n", rbf_exp64(1e-12), rbf_exp64(0.0), rbf_exp64(-0.5), (double)exp2l(-0.5L/64));
  return 0;}
