/* walker_step_check.c -- CPU restatement of the branch-free, six-stage phase-1 walker step (stochasticsim_b200/csrc/spike_chain.cuh, walk_seg)
 * checked against the plain per-locus loop it stands for (selectMutantAllele / randomNum, stochasticSpike.c:283-302, 338-360):
 * per covered locus, draw until the pick is neither rejected by randomNum nor equal to the reference class; a locus whose
 * reference base is not one of GCAT takes the first draw randomNum accepts.  Random planes, random stretches, 20000 rounds.
 * Test infrastructure: compiled and run by tests/test_walker_step.py; the expressions below must stay in step with the kernel. */
#include <stdio.h>
#include <stdlib.h>
#include <stdint.h>
#include <string.h>
#define NK 4096
#define NG 1024
static uint32_t e0[NK/32+4], e1[NK/32+4], ej[NK/32+4], c0[NG/32+4], c1[NG/32+4], cx[NG/32+4];
static int ecls[NK], ccls[NG];
static uint32_t fsr(uint32_t lo, uint32_t hi, uint32_t s){ s&=31; return s? (lo>>s)|(hi<<(32-s)) : lo; }
#define STAGES 6
static uint32_t shl(uint32_t v, uint32_t s){ return s>=32u ? 0u : v<<s; }      /* shl.b32 clamps */
static int walk_seg(uint32_t *pgr, uint32_t gto, uint32_t *pkr){
  uint32_t gr=*pgr,kr=*pkr;
  while(gr+32u<=gto){                       /* a full window of loci: STAGES conflicts per step */
    uint32_t a=kr&31,b=gr&31,we=kr>>5,wc=gr>>5;
    if (we+2 > NK/32) return 1;
    uint32_t E0=fsr(e0[we],e0[we+1],a),E1=fsr(e1[we],e1[we+1],a),EJ=fsr(ej[we],ej[we+1],a);
    uint32_t C0=fsr(c0[wc],c0[wc+1],b),C1=fsr(c1[wc],c1[wc+1],b),CX=fsr(cx[wc],cx[wc+1],b);
    uint32_t done=0,nonend=0,skew=0;
    for(int s=0;s<STAGES;s++){
      uint32_t s0=shl(C0,skew),s1=shl(C1,skew),sx=shl(CX,skew);
      uint32_t term=(((E0^s0)|(E1^s1)|sx)&~EJ)|done;
      uint32_t z=~term&(term+1u);
      uint32_t m0=(s0&z)?0xffffffffu:0u,m1=(s1&z)?0xffffffffu:0u,mx=(sx&z)?0xffffffffu:0u;
      uint32_t ends=((E0^m0)|(E1^m1)|mx)&~EJ;
      uint32_t above=ends&~(z|(z-1u));
      uint32_t y=above&(0u-above);
      nonend|=y-z;
      done=y|(y-1u);
      skew=(uint32_t)__builtin_popcount(nonend);
    }
    uint32_t d=(uint32_t)__builtin_popcount(done);
    kr+=d; gr+=d-(uint32_t)__builtin_popcount(nonend);
  }
  while(gr<gto){
    uint32_t a=kr&31,b=gr&31,we=kr>>5,wc=gr>>5;
    if (we+2 > NK/32) return 1;
    uint32_t E0=fsr(e0[we],e0[we+1],a),E1=fsr(e1[we],e1[we+1],a),EJ=fsr(ej[we],ej[we+1],a);
    uint32_t C0=fsr(c0[wc],c0[wc+1],b),C1=fsr(c1[wc],c1[wc+1],b),CX=fsr(cx[wc],cx[wc+1],b);
    uint32_t n = 32u < gto-gr ? 32u : gto-gr;
    uint32_t beyond = n>=32u ? 0u : (0xffffffffu<<n);
    uint32_t term=(((E0^C0)|(E1^C1)|CX)&~EJ)|beyond;
    uint32_t z=~term&(term+1u);
    uint32_t m0=(C0&z)?0xffffffffu:0u,m1=(C1&z)?0xffffffffu:0u,mx=(CX&z)?0xffffffffu:0u;
    uint32_t ends=((E0^m0)|(E1^m1)|mx)&~EJ;
    uint32_t above=ends&~(z|(z-1u));
    uint32_t y=above&(0u-above);
    uint32_t t=(uint32_t)__builtin_popcount(z-1u); if (t>n) t=n;
    gr += t + (y?1u:0u);
    kr += y ? (uint32_t)__builtin_popcount(y-1u)+1u : (z?32u:n);
  }
  *pgr=gr;*pkr=kr;return 0;
}
static int naive(uint32_t *pgr, uint32_t gto, uint32_t *pkr){
  uint32_t gr=*pgr,kr=*pkr;
  while(gr<gto){
    for(;;){ if(kr>=NK-64) return 1; int e=ecls[kr++]; if(e==5) continue; if(ccls[gr]==4) break; if(e!=ccls[gr]) break; }
    gr++;
  }
  *pgr=gr;*pkr=kr;return 0;
}
int main(){
  srand(1);
  for(int it=0;it<20000;it++){
    memset(e0,0,sizeof e0);memset(e1,0,sizeof e1);memset(ej,0,sizeof ej);memset(c0,0,sizeof c0);memset(c1,0,sizeof c1);memset(cx,0,sizeof cx);
    int pj = rand()%3==0? 30: 2;
    for(int k=0;k<NK;k++){ int r=rand()%100; int e = r<pj?5:rand()%4; ecls[k]=e; if(e==5){ej[k>>5]|=1u<<(k&31); int z=rand()%4; if(z&1)e0[k>>5]|=1u<<(k&31); if(z&2)e1[k>>5]|=1u<<(k&31);} else { if(e&1)e0[k>>5]|=1u<<(k&31); if(e&2)e1[k>>5]|=1u<<(k&31);} }
    int px = rand()%4==0? 20:1; int same = rand()%5==0;
    for(int g=0;g<NG;g++){ int c = (rand()%100<px)?4:(same?1:rand()%4); ccls[g]=c; if(c==4){cx[g>>5]|=1u<<(g&31); int z=rand()%4; if(z&1)c0[g>>5]|=1u<<(g&31); if(z&2)c1[g>>5]|=1u<<(g&31);} else { if(c&1)c0[g>>5]|=1u<<(g&31); if(c&2)c1[g>>5]|=1u<<(g&31);} }
    uint32_t g0=rand()%200, gto=g0+rand()%700, k0=rand()%300;
    uint32_t g1=g0,k1=k0,g2=g0,k2=k0;
    int r1=walk_seg(&g1,gto,&k1), r2=naive(&g2,gto,&k2);
    if(r2==0 && (r1!=0 || g1!=g2||k1!=k2)){printf("MISMATCH it=%d g %u %u k %u %u r1=%d\n",it,g1,g2,k1,k2,r1);return 1;}
  }
  printf("ok\n");return 0;
}
