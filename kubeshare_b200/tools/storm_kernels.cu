// storm_kernels.cu -- the synthetic client kernels of the benchmark: an empty kernel for the
// launch storm (configs 1-2), a ~N-microsecond spin kernel for the bursty trace (config 3), and an
// MNIST-shaped 3x3 convolution (config 5).  Plain CUDA: these are the APPLICATION's kernels, not the hook's.
#include <stdint.h>
extern "C" {
__global__ void noop() {}

__global__ void spin(unsigned long long ns) {
  unsigned long long t0, t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t0));
  do {
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  } while (t - t0 < ns);
}

// Independent truth for the accounting tests: the kernel spins for `ns` and folds its own %globaltimer start/end into
// the bounds of its burst (bounds[2b] = earliest start, bounds[2b+1] = latest end), so the busy span of every burst is
// known from the device's own clock without any event.
__global__ void spin_stamp(unsigned long long ns, unsigned long long* __restrict__ bounds, unsigned burst) {
  unsigned long long t0, t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t0));
  do {
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  } while (t - t0 < ns);
  if (threadIdx.x == 0) {
    atomicMin(bounds + 2 * burst, t0);
    atomicMax(bounds + 2 * burst + 1, t);
  }
}

// out[n][co][y][x] = relu(sum_{ci,ky,kx} in[n][ci][y+ky-1][x+kx-1] * w[co][ci][ky][kx]); 28x28 images
__global__ void conv3x3(const float* __restrict__ in, const float* __restrict__ w, float* __restrict__ out,
                        int cin, int cout) {
  int x = threadIdx.x, y = threadIdx.y, co = blockIdx.x, n = blockIdx.y;
  float acc = 0.f;
  for (int ci = 0; ci < cin; ci++)
    for (int ky = 0; ky < 3; ky++)
      for (int kx = 0; kx < 3; kx++) {
        int yy = y + ky - 1, xx = x + kx - 1;
        if (yy >= 0 && yy < 28 && xx >= 0 && xx < 28)
          acc += in[((n * cin + ci) * 28 + yy) * 28 + xx] * w[((co * cin + ci) * 3 + ky) * 3 + kx];
      }
  out[((n * cout + co) * 28 + y) * 28 + x] = acc > 0.f ? acc : 0.f;
}
}
