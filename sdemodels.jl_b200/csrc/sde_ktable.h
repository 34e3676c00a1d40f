/* sde_ktable.h — launch handles of one compiled model variant (AOT: addresses of __global__ functions;
 * NVRTC: cudaKernel_t handles).  Index [scheme][precision]: scheme 0 = EulerMaruyama, 1 = Milstein, 2 = RungeKutta;
 * precision 0 = f64, 1 = f32.  A null entry means "not defined for this model" (Milstein / RungeKutta with M != 1). */
#ifndef SDE_KTABLE_H
#define SDE_KTABLE_H
typedef struct {
  const void *sample[3][2];
  const void *mlmc[3][2];
  const void *simsoa[3][2];
  const void *simref[3][2];
  const void *simw[3][2];
  const void *coef;
} sde_ktable;
#endif
