/* sde_ktable.h — launch handles of one compiled model variant (AOT: addresses of __global__ functions;
 * NVRTC: cudaKernel_t handles).  Index [scheme][slot]: scheme 0 = EulerMaruyama, 1 = Milstein, 2 = RungeKutta, 3 = Exact;
 * slot 0 = f64 on normal stream v1, 1 = f32, 2 = f64 on normal stream v2 (kernels without noise: the same handle as slot 0).  A null entry means "not defined for this model" (Milstein / RungeKutta with M != 1). */
#ifndef SDE_KTABLE_H
#define SDE_KTABLE_H
typedef struct {
  const void *sample[4][3];
  const void *sample1[4][3]; /* one path per thread, 512-thread CTAs: launches too small to fill the GPU with SDE_PPT paths per thread */
  const void *mlmc[4][3];
  const void *simsoa[4][3];
  const void *simref[4][3];
  const void *simrt[4][3]; /* reference-order trajectories with TMA tensor stores */
  const void *simw[4][3];
  const void *simwt[4][3]; /* Wiener-driven on TMA tensor tiles (null: not applicable) */
  const void *logt[2]; /* [precision] */
  const void *logtt[2]; /* logtransition on TMA tensor loads (null: not applicable) */
  const void *coef;
} sde_ktable;
#endif
