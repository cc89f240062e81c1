/* abi_driver.c -- a plain C caller of libnsgym_b200.so: the README quickstart of the reference
 * (NS CartPole: masspole IncrementUpdate(k = 0.1) on a ContinuousScheduler, gravity RandomWalk on
 * PeriodicScheduler(3); README.md:107-117) stepped through the C ABI alone -- no Python, no torch.
 * Device memory comes from the CUDA runtime.  Prints one line per check and exits non-zero on the
 * first failure.  Built and run by tests/test_c_abi.py.
 *
 *   gcc -std=c99 -I include -I $CUDA/include tests/c/abi_driver.c -o abi_driver \
 *       -L ns_gym_b200/_lib -lnsgym_b200 -L $CUDA/lib64 -lcudart -lm
 */
#include <cuda_runtime_api.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "nsgym_b200.h"

#define CHECK(cond, ...)                                  \
  do {                                                    \
    if (!(cond)) {                                        \
      fprintf(stderr, "FAIL %s:%d: ", __FILE__, __LINE__); \
      fprintf(stderr, __VA_ARGS__);                       \
      fprintf(stderr, " (%s)\n", nsgym_last_error());     \
      return 1;                                           \
    }                                                     \
  } while (0)

int main(void) {
  const int64_t n = 10000;
  const int steps = 25;
  NsgymSpec spec;
  memset(&spec, 0, sizeof spec);
  spec.abi_version = NSGYM_ABI_VERSION;
  spec.env_kind = NSGYM_ENV_CARTPOLE;
  spec.precision = NSGYM_F64;
  spec.autoreset = NSGYM_AUTORESET_NONE;
  spec.n_envs = n;
  spec.seed = 7;
  spec.max_episode_steps = 500;
  spec.n_slots = 2;
  /* theta order of CartPole: gravity masscart masspole force_mag tau length */
  const double defaults[6] = {9.8, 1.0, 0.1, 10.0, 0.02, 0.5};
  for (int i = 0; i < 6; ++i) spec.theta_init[i][0] = defaults[i];
  spec.slots[0].theta_index = 2;                       /* masspole */
  spec.slots[0].sched_op = NSGYM_SCHED_CONTINUOUS;
  spec.slots[0].upd_op = NSGYM_UPD_ADD;
  spec.slots[0].uf[0] = 0.1;
  spec.slots[0].constraint = NSGYM_CONS_REJECT_LE0;
  spec.slots[1].theta_index = 0;                       /* gravity */
  spec.slots[1].sched_op = NSGYM_SCHED_PERIODIC;
  spec.slots[1].si[0] = 3;
  spec.slots[1].upd_op = NSGYM_UPD_RW;
  spec.slots[1].uf[1] = 0.0;                           /* mu */
  spec.slots[1].uf[2] = 1.0;                           /* sigma */
  spec.slots[1].constraint = NSGYM_CONS_REJECT_LT0;
  for (int j = 0; j < 2; ++j) {
    spec.slots[j].end = INT32_MAX;
    spec.slots[j].partner_slot = -1;
    spec.slots[j].istate_plane = -1;
  }

  CHECK(nsgym_abi_version() == NSGYM_ABI_VERSION, "ABI version");
  CHECK(nsgym_sizeof(1) == sizeof(NsgymSpec) && nsgym_sizeof(0) == sizeof(NsgymSlot), "struct sizes");
  NsgymHandle* h = NULL;
  CHECK(nsgym_create(&spec, &h) == 0, "nsgym_create");
  NsgymLayout lay;
  CHECK(nsgym_layout(h, 1, 1, &lay) == 0, "nsgym_layout");
  CHECK(lay.state == (size_t)n * 4 * 8 && lay.theta == (size_t)n * 2 * 8 && lay.theta_planes == 2, "layout");
  printf("layout ok: %.0f algorithmic bytes per env-step\n", lay.bytes_per_step);

  NsgymBuffers b;
  memset(&b, 0, sizeof b);
  void** slots[] = {&b.d_state, &b.d_theta, (void**)&b.d_t, &b.d_action, (void**)&b.d_reward, (void**)&b.d_flags,
                    (void**)&b.d_change, &b.d_delta, (void**)&b.d_obs};
  const size_t bytes[] = {lay.state, lay.theta, lay.t, lay.action, lay.reward, lay.flags, lay.change, lay.delta, lay.obs};
  for (int k = 0; k < 9; ++k) {
    CHECK(cudaMalloc(slots[k], bytes[k]) == cudaSuccess, "cudaMalloc");
    CHECK(cudaMemset(*slots[k], 0, bytes[k]) == cudaSuccess, "cudaMemset");
  }
  CHECK(nsgym_bind(h, &b) == 0, "nsgym_bind");
  CHECK(nsgym_step(h, NULL, NULL, NULL, 0, NULL) == -4, "a step before the first reset must be refused");
  CHECK(nsgym_reset(h, NULL, NULL, NULL) == 0, "nsgym_reset");

  int32_t* actions = (int32_t*)malloc(sizeof(int32_t) * (size_t)n);
  for (int64_t i = 0; i < n; ++i) actions[i] = (int32_t)(i & 1);
  CHECK(cudaMemcpy(b.d_action, actions, sizeof(int32_t) * (size_t)n, cudaMemcpyHostToDevice) == cudaSuccess, "H2D");
  for (int k = 0; k < steps; ++k) CHECK(nsgym_step(h, NULL, NULL, NULL, 0, NULL) == 0, "nsgym_step %d", k);
  CHECK(cudaDeviceSynchronize() == cudaSuccess, "sync");
  CHECK(nsgym_launch_count(h) == steps + 1, "launch count");

  double* theta = (double*)malloc(lay.theta);
  int32_t* t = (int32_t*)malloc(lay.t);
  double* state = (double*)malloc(lay.state);
  uint8_t* change = (uint8_t*)malloc(lay.change);
  CHECK(cudaMemcpy(theta, b.d_theta, lay.theta, cudaMemcpyDeviceToHost) == cudaSuccess, "D2H");
  CHECK(cudaMemcpy(t, b.d_t, lay.t, cudaMemcpyDeviceToHost) == cudaSuccess, "D2H");
  CHECK(cudaMemcpy(state, b.d_state, lay.state, cudaMemcpyDeviceToHost) == cudaSuccess, "D2H");
  CHECK(cudaMemcpy(change, b.d_change, lay.change, cudaMemcpyDeviceToHost) == cudaSuccess, "D2H");
  double gsum = 0.0, gsq = 0.0;
  int rejected = 0;
  for (int64_t i = 0; i < n; ++i) {
    /* masspole: 0.1 + 25 * 0.1 (theta advanced with the pre-increment t, every step) */
    CHECK(fabs(theta[i] - (0.1 + 0.1 * steps)) < 1e-12, "masspole of env %lld = %.17g", (long long)i, theta[i]);
    CHECK((t[i] & 0x0FFFFFFF) == steps, "relative_time of env %lld", (long long)i);
    /* last step had t = 24: the period-3 scheduler fired, masspole always fires; a gravity candidate
     * below zero is rejected by the constraint checker (classic_control.py:208-235): flag 0 */
    CHECK(change[i] == 3 || change[i] == 1, "change mask of env %lld = %d", (long long)i, change[i]);
    rejected += change[i] == 1;
    const double g = theta[n + i] - 9.8;      /* sum of 9 standard normals (t = 0, 3, .., 24) */
    gsum += g;
    gsq += g * g;
    CHECK(isfinite(state[4 * i]) && isfinite(state[4 * i + 2]), "state of env %lld", (long long)i);
  }
  const double mean = gsum / n, var = gsq / n - mean * mean;
  printf("gravity random walk after 9 fires: mean %+.4f (0 expected), variance %.3f (9 expected)\n", mean, var);
  CHECK(fabs(mean) < 0.15 && fabs(var - 9.0) < 0.6, "random-walk moments");
  CHECK(rejected < n / 100, "%d gravity updates rejected", rejected);

  /* error behaviour: invalid specs are refused with a message, nothing is thrown */
  NsgymSpec bad = spec;
  bad.slots[1].theta_index = 2;
  NsgymHandle* h2 = NULL;
  CHECK(nsgym_create(&bad, &h2) == -1 && h2 == NULL && strstr(nsgym_last_error(), "bound twice"), "duplicate parameter");
  bad = spec;
  bad.abi_version = 99;
  CHECK(nsgym_create(&bad, &h2) == -1 && strstr(nsgym_last_error(), "ABI version"), "ABI mismatch");

  nsgym_destroy(h);
  printf("C ABI driver OK\n");
  return 0;
}
