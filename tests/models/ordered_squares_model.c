// CPU model of ordered_squares_* (qr_exact_kernels.cuh) against the plain chain
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
typedef unsigned long long u64;
typedef struct { u64 d0, d1; } SqPair;
static u64 bitsof(double x){u64 b;memcpy(&b,&x,8);return b;}
static double frombits(u64 b){double x;memcpy(&x,&b,8);return x;}
static void sq_addend(double v,int fused,u64*hi,u64*lo,int*x){
  if(fused){u64 bits=bitsof(v)&0x7fffffffffffffffull;int ef=(int)(bits>>52);u64 m=ef?((bits&0xfffffffffffffull)|0x10000000000000ull):bits;int e=ef?ef-1075:-1074;unsigned __int128 p=(unsigned __int128)m*m;*lo=(u64)p;*hi=(u64)(p>>64);*x=2*e;}
  else{volatile double a=v*v;u64 bits=bitsof(a);int ef=(int)(bits>>52);*lo=ef?((bits&0xfffffffffffffull)|0x10000000000000ull):bits;*hi=0;*x=ef?ef-1075:-1074;}
}
static int sq_element(double v,int fused,int E,SqPair*out){
  u64 hi,lo;int x;sq_addend(v,fused,&hi,&lo,&x);out->d0=out->d1=0;if((hi|lo)==0)return 1;
  int sh=(E-52)-x;
  if(sh<=0){if(hi!=0||sh<-11||(lo>>(53+sh))!=0)return 0;out->d0=out->d1=lo<<(-sh);return 1;}
  if(sh>=128)return 1;
  u64 q,qh,half,sticky;
  if(sh>=64){int s2=sh-64;q=s2?hi>>s2:hi;qh=0;if(s2==0){half=lo>>63;sticky=lo<<1;}else{half=(hi>>(s2-1))&1;sticky=(s2>1?hi<<(65-s2):0)|lo;}}
  else{q=(lo>>sh)|(hi<<(64-sh));qh=hi>>sh;half=(lo>>(sh-1))&1;sticky=sh>1?lo<<(65-sh):0;}
  if(qh!=0||(q>>53)!=0)return 0;
  if(half&&sticky){out->d0=out->d1=q+1;}else if(half){out->d0=q+(q&1);out->d1=q+((q+1)&1);}else{out->d0=out->d1=q;}
  return 1;
}
static SqPair compose(SqPair a,SqPair b){SqPair r;r.d0=a.d0+((a.d0&1)?b.d1:b.d0);r.d1=a.d1+(((a.d1+1)&1)?b.d1:b.d0);return r;}
static double step(double acc,double v,int fused){ if(fused) return fma(v,v,acc); volatile double p=v*v; return acc+p; }
#define CH 256
int main(int argc,char**argv){
  int n=argc>1?atoi(argv[1]):1000000; int mode=argc>2?atoi(argv[2]):0; srand(argc>3?atoi(argv[3]):1);
  double*v=malloc(n*sizeof(double));
  for(int i=0;i<n;i++){double u=(rand()+0.5)/RAND_MAX; double s=(rand()&1)?1:-1;
    if(mode==0) v[i]=s*u*1e-3; else if(mode==1) v[i]=s*ldexp(u,(rand()%40)-20); else if(mode==2) v[i]=(rand()%4==0)?0.0:s*ldexp((double)(rand()%8+1),-3);  /* many exact ties */
    else if(mode==3) v[i]=s*ldexp(u,-(i%1000)); else v[i]= i==0 ? 741456.0 : s*ldexp((double)(rand()%15+1),-7); }
  for(int fused=0;fused<2;fused++){
    double seq=0;for(int i=0;i<n;i++)seq=step(seq,v[i],fused);
    int nch=(n+CH-1)/CH;double acc=0;double approx=0;int replay=0;
    for(int c=0;c<nch;c++){
      int i0=c*CH,i1=i0+CH<n?i0+CH:n;
      u64 sb=bitsof(approx);int ef=(int)(sb>>52);int E=(ef>=64&&ef<2047)?ef-1023:-99999;
      double cs=0;for(int i=i0;i<i1;i++)cs+=v[i]*v[i];
      int ok=E!=-99999;SqPair a={0,0};
      if(ok){for(int i=i0;i<i1;i++){SqPair e;ok=sq_element(v[i],fused,E,&e)&&ok;a=compose(a,e);} if((a.d0>>53)||(a.d1>>53))ok=0;}
      u64 bits=bitsof(acc);int done=0;
      if(ok&&(int)(bits>>52)-1023==E){u64 m=(bits&0xfffffffffffffull)|0x10000000000000ull;u64 m2=m+((m&1)?a.d1:a.d0);if((m2>>53)==0){acc=frombits((bits&0x7ff0000000000000ull)|(m2&0xfffffffffffffull));done=1;}}
      if(!done){replay++;for(int i=i0;i<i1;i++)acc=step(acc,v[i],fused);}
      approx+=cs;
    }
    printf("n=%d mode=%d fused=%d seq=%.17g par=%.17g %s replayed=%d/%d\n",n,mode,fused,seq,acc,bitsof(seq)==bitsof(acc)?"EQUAL":"DIFFER",replay,nch);
  }
  return 0;
}
