// Host-side finite-difference check of the device cost functions (covins_b200/csrc/ba_math.cuh is __host__ __device__):
// reprojection, between and IMU Jacobians in the local parametrisation.  Built and run by tests/test_ba_math_host.py.
#define __host__
#define __device__
#include "../../covins_b200/csrc/ba_math.cuh"
#include <cstdio>
#include <cstdlib>
#include <cstring>
using namespace bam;
double rnd(){return rand()/(double)RAND_MAX*2-1;}
int main(){
  srand(1);
  double pose[7]={0.1,-0.2,0.3,0.9,1,2,3}; {double n=sqrt(pose[0]*pose[0]+pose[1]*pose[1]+pose[2]*pose[2]+pose[3]*pose[3]); for(int i=0;i<4;i++)pose[i]/=n;}
  double extr[7]={0.5,-0.5,0.5,0.5,0.02,-0.06,0.01};
  double intr[4]={458.654,457.296,367.215,248.375}, dist[4]={-0.28340811,0.07395907,0.00019359,1.76187114e-05};
  // all six GlobalEuclideanReprError<Camera, Distortion> instantiations (optimization_be.cpp:186-231)
  const double dists[3][4]={{-0.28340811,0.07395907,0.00019359,1.76187114e-05},{-0.013,0.02,-0.012,0.002},{0.93,0,0,0}};
  double lm[3];
  for(int cam=0;cam<2;cam++) for(int dm=0;dm<3;dm++){
    CamModel cm{cam,dm,cam?1.2:0.0};
    memcpy(dist,dists[dm],32);
    for(;;){ lm[0]=1+3*rnd(); lm[1]=2+3*rnd(); lm[2]=3+3*rnd(); double r[2],Jp[12],Jl[6]; if(reproj(pose,extr,intr,dist,cm,lm,300,200,2.0,r,Jp,Jl,true)) { if (fabs(r[0])<200&&fabs(r[1])<200) break;} }
    double r[2],Jp[12],Jl[6]; reproj(pose,extr,intr,dist,cm,lm,300,200,2.0,r,Jp,Jl,true);
    double eps=1e-6, maxe=0;
    for(int c=0;c<6;c++){ double d[6]={0,0,0,0,0,0}; d[c]=eps; double pp[7],pm[7]; pose_plus(pose,d,pp); d[c]=-eps; pose_plus(pose,d,pm);
      double rp[2],rm[2]; reproj(pp,extr,intr,dist,cm,lm,300,200,2.0,rp,0,0,false); reproj(pm,extr,intr,dist,cm,lm,300,200,2.0,rm,0,0,false);
      for(int a=0;a<2;a++){ double fd=(rp[a]-rm[a])/(2*eps); maxe=fmax(maxe,fabs(fd-Jp[6*a+c])); } }
    for(int c=0;c<3;c++){ double lp[3],lmn[3]; memcpy(lp,lm,24); memcpy(lmn,lm,24); lp[c]+=eps; lmn[c]-=eps; double rp[2],rm[2];
      reproj(pose,extr,intr,dist,cm,lp,300,200,2.0,rp,0,0,false); reproj(pose,extr,intr,dist,cm,lmn,300,200,2.0,rm,0,0,false);
      for(int a=0;a<2;a++){ double fd=(rp[a]-rm[a])/(2*eps); maxe=fmax(maxe,fabs(fd-Jl[3*a+c])); } }
    printf("reproj cam %d dist %d r=(%g,%g) max jac err %g (|J| ~ %g)\n", cam,dm,r[0],r[1],maxe,fabs(Jp[0]));
  }
  double eps=1e-6, maxe=0;
  // between
  double pose2[7]={-0.3,0.1,0.2,0.8,2,1,4}; {double n=sqrt(pose2[0]*pose2[0]+pose2[1]*pose2[1]+pose2[2]*pose2[2]+pose2[3]*pose2[3]); for(int i=0;i<4;i++)pose2[i]/=n;}
  double qm[4]={0.05,0.1,-0.1,0.98}; {double n=sqrt(qm[0]*qm[0]+qm[1]*qm[1]+qm[2]*qm[2]+qm[3]*qm[3]); for(int i=0;i<4;i++)qm[i]/=n;}
  double tm[3]={0.3,0.2,0.1}, S[36]; for(int i=0;i<36;i++) S[i]=rnd(); 
  double e[6],J[72]; between(pose,pose2,qm,tm,S,e,J,true); maxe=0;
  for(int c=0;c<12;c++){ double d[6]={0,0,0,0,0,0}; d[c%6]=eps; double pp[7],pm[7]; const double* base=c<6?pose:pose2; pose_plus(base,d,pp); d[c%6]=-eps; pose_plus(base,d,pm);
    double ep[6],em[6]; if(c<6){between(pp,pose2,qm,tm,S,ep,0,false);between(pm,pose2,qm,tm,S,em,0,false);} else {between(pose,pp,qm,tm,S,ep,0,false);between(pose,pm,qm,tm,S,em,0,false);}
    for(int a=0;a<6;a++){ double fd=(ep[a]-em[a])/(2*eps); maxe=fmax(maxe,fabs(fd-J[12*a+c])); } }
  printf("between max jac err %g\n",maxe);
  // imu
  ImuPre P; memset(&P,0,sizeof(P)); P.T=0.25; for(int i=0;i<3;i++){P.alpha[i]=0.1*rnd();P.beta[i]=rnd();P.ba[i]=0.01*rnd();P.bg[i]=0.01*rnd();}
  Q4 g=qexp(V3{0.1,-0.05,0.07}); P.gamma[0]=g.x;P.gamma[1]=g.y;P.gamma[2]=g.z;P.gamma[3]=g.w;
  for(int i=0;i<9;i++){P.dp_dba[i]=0.1*rnd();P.dp_dbg[i]=0.1*rnd();P.dq_dbg[i]=rnd();P.dv_dba[i]=rnd();P.dv_dbg[i]=rnd();}
  double sbi[9],sbj[9]; for(int i=0;i<9;i++){sbi[i]=0.3*rnd();sbj[i]=0.3*rnd();}
  double rr[15],Jr[450]; imu_raw(pose,sbi,pose2,sbj,P,9.81,rr,Jr); maxe=0; double worst=-1; int wc=-1,wa=-1;
  for(int c=0;c<30;c++){ double pi[7],pj[7],si[9],sj[9],pi2[7],pj2[7],si2[9],sj2[9]; memcpy(pi,pose,56);memcpy(pj,pose2,56);memcpy(si,sbi,72);memcpy(sj,sbj,72);
    memcpy(pi2,pose,56);memcpy(pj2,pose2,56);memcpy(si2,sbi,72);memcpy(sj2,sbj,72);
    double d[6]={0,0,0,0,0,0};
    if(c<6){d[c]=eps;pose_plus(pose,d,pi);d[c]=-eps;pose_plus(pose,d,pi2);} else if(c<15){si[c-6]+=eps;si2[c-6]-=eps;} else if(c<21){d[c-15]=eps;pose_plus(pose2,d,pj);d[c-15]=-eps;pose_plus(pose2,d,pj2);} else {sj[c-21]+=eps;sj2[c-21]-=eps;}
    double rp[15],rm[15]; imu_raw(pi,si,pj,sj,P,9.81,rp,0); imu_raw(pi2,si2,pj2,sj2,P,9.81,rm,0);
    for(int a=0;a<15;a++){ double fd=(rp[a]-rm[a])/(2*eps); double er=fabs(fd-Jr[30*a+c]); if(er>worst){worst=er;wc=c;wa=a;} } }
  printf("imu max jac err %g at row %d col %d\n",worst,wa,wc);
  // the column-wise evaluation used by the warp-per-factor kernel must reproduce imu_raw exactly
  double cmax=0, rr2[15];
  for(int c=0;c<30;c++){ double col[15]; imu_raw_column(pose,sbi,pose2,sbj,P,9.81,c,col,c==0?rr2:nullptr); for(int a=0;a<15;a++) cmax=fmax(cmax,fabs(col[a]-Jr[30*a+c])); }
  for(int a=0;a<15;a++) cmax=fmax(cmax,fabs(rr2[a]-rr[a]));
  printf("imu column-wise vs full: max abs diff %g\n",cmax);
  if(cmax!=0.0) return 1;
}
