/* Host-side validation of the part-pair ranking used by k_pair_eval (csrc/pair_kernels.cuh, pair_three_both, PAIR_HMIN_RANKED):
 * distance_three_circles (reference core/distance.py:55-105) takes the minimum of h = hypot(x, y) - (r_i + r_j) over nine part
 * pairs with a strict '<' in a fixed order.  The kernel ranks the nine in fp32 and runs the exact hypot only for those within a
 * margin of the fp32 minimum.  exact() is the reference's loop, fast() the kernel's selection with the same constants; they must
 * agree BIT FOR BIT (h_min, the winning pair, x, y, d) -- random pairs, pairs far from the origin, coincident agents, zero radii
 * and configurations with exact ties.  Usage: validate_hmin_ranking [pairs per mode, default 3000000]; exit status 1 on mismatch.
 * Test infrastructure (tests/test_hmin_ranking_host.py); nothing in the product links it. */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
typedef struct { double x[3], y[3], r[3]; } A3;
static void exact(const A3*I,const A3*J,double*h_min,double*sx,double*sy,double*sd,int*im,int*jm){
  *h_min=NAN;*sx=*sy=*sd=0;*im=*jm=0;
  for(int pi=0;pi<3;pi++)for(int pj=0;pj<3;pj++){double x=I->x[pi]-J->x[pj],y=I->y[pi]-J->y[pj];double d=hypot(x,y);double h=d-(I->r[pi]+J->r[pj]);
    if(h<*h_min||isnan(*h_min)){*h_min=h;*sx=x;*sy=y;*sd=d;*im=pi;*jm=pj;}}
}
static int ncand_tot=0;
static void fast(const A3*I,const A3*J,double*h_min,double*sx,double*sy,double*sd,int*im,int*jm){
  float rif[3]={(float)I->r[0],(float)I->r[1],(float)I->r[2]},rjf[3]={(float)J->r[0],(float)J->r[1],(float)J->r[2]};
  float ha[9],lo=3.0e38f,dhi=0,chk=0;
  for(int k=0;k<9;k++){int pi=k/3,pj=k%3;double x=I->x[pi]-J->x[pj],y=I->y[pi]-J->y[pj];float df=sqrtf((float)fma(x,x,y*y));ha[k]=df-(rif[pi]+rjf[pj]);lo=fminf(lo,ha[k]);dhi=fmaxf(dhi,df);chk+=ha[k];}
  float thr=lo+2.0f*(1e-6f*(dhi+rif[0]+rif[1]+rjf[0]+rjf[1])+1e-15f);
  unsigned cand=0;for(int k=0;k<9;k++)if(ha[k]<=thr)cand|=1u<<k;
  if(!(fabsf(chk)<1e30f))cand=0x1ff;
  ncand_tot+=__builtin_popcount(cand);
  *h_min=NAN;*sx=*sy=*sd=0;*im=*jm=0;
  while(cand){int k=__builtin_ffs(cand)-1;cand&=cand-1;int pi=(k>=3)+(k>=6),pj=k-3*pi;double x=I->x[pi]-J->x[pj],y=I->y[pi]-J->y[pj];double d=hypot(x,y);double h=d-(I->r[pi]+J->r[pj]);
    if(h<*h_min||isnan(*h_min)){*h_min=h;*sx=x;*sy=y;*sd=d;*im=pi;*jm=pj;}}
}
static double U(){return rand()/(double)RAND_MAX;}
static void mk(A3*a,double cx,double cy,double phi,double rt,double rs,double rts){a->x[0]=cx;a->y[0]=cy;double ox=rts*sin(phi),oy=-rts*cos(phi);a->x[1]=cx-ox;a->y[1]=cy-oy;a->x[2]=cx+ox;a->y[2]=cy+oy;a->r[0]=rt;a->r[1]=rs;a->r[2]=rs;}
int main(int argc,char**argv){long bad=0,n=0,iters=argc>1?atol(argv[1]):3000000;srand(1);
  for(int mode=0;mode<8;mode++)for(long it=0;it<iters;it++){A3 I,J;double base=(mode==3)?1e5*(U()-0.5):(mode==4?1e7:0);
    double rt=0.1+0.1*U(),rs=0.05+0.08*U(),rts=0.1+0.1*U();double rt2=rt,rs2=rs,rts2=rts;if(mode!=1&&mode!=5){rt2=0.1+0.1*U();rs2=0.05+0.08*U();rts2=0.1+0.1*U();}
    double phi=(U()-0.5)*6.283,phi2=(U()-0.5)*6.283;double dist=mode==2?1e-9*U():(mode==6?1e-3*U():4*U());double ang=(U()-0.5)*6.283;
    if(mode==1||mode==5){ /* symmetric / aligned: exact ties */ phi2=phi; if(mode==5){phi=0.5*3.141592653589793*(rand()%4);phi2=phi+3.141592653589793*(rand()%2);ang=0.5*3.141592653589793*(rand()%4);dist=0.25*(rand()%16);} }
    if(mode==7){rt=rt2=rs=rs2=0.0;rts=rts2=1e-3*U();dist=1e-4*U();}
    mk(&I,base+0.0,base+0.0,phi,rt,rs,rts);mk(&J,base+dist*cos(ang),base+dist*sin(ang),phi2,rt2,rs2,rts2);
    double h1,sx1,sy1,sd1,h2,sx2,sy2,sd2;int i1,j1,i2,j2;exact(&I,&J,&h1,&sx1,&sy1,&sd1,&i1,&j1);fast(&I,&J,&h2,&sx2,&sy2,&sd2,&i2,&j2);n++;
    if(memcmp(&h1,&h2,8)||memcmp(&sx1,&sx2,8)||memcmp(&sy1,&sy2,8)||memcmp(&sd1,&sd2,8)||i1!=i2||j1!=j2){if(bad<5)printf("MISMATCH mode %d: %g %g (%d,%d) vs (%d,%d)\n",mode,h1,h2,i1,j1,i2,j2);bad++;}}
  printf("pairs %ld mismatches %ld mean candidates %.3f\n",n,bad,ncand_tot/(double)n);return bad!=0;}
