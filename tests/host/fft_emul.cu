// Host emulation of the shared-memory FFT index arithmetic in remfx_b200/csrc/fft.cuh.
// Build:  nvcc -O2 -o /tmp/fft_emul tests/host/fft_emul.cu   (runs on the CPU; no GPU needed)
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../../remfx_b200/csrc/fft.cuh"
using namespace rfx;

static int run(int n_fft) {
  const int NC = n_fft / 2;
  int log2nc = 0;
  while ((1 << log2nc) < NC) ++log2nc;
  std::vector<float2> tw(n_fft);
  for (int m = 0; m < n_fft; ++m) tw[m] = make_float2((float)cos(-2.0 * M_PI * m / n_fft), (float)sin(-2.0 * M_PI * m / n_fft));
  std::vector<float> x(n_fft);
  srand(n_fft);
  for (auto& v : x) v = (float)rand() / RAND_MAX - 0.5f;
  std::vector<float2> a(NC), b(NC);
  for (int n = 0; n < NC; ++n) a[n] = make_float2(x[2 * n], x[2 * n + 1]);
  float2 *in = a.data(), *out = b.data();
  for (int Ns = 1; Ns * 4 <= NC; Ns *= 4) {
    for (int j = 0; j < NC / 4; ++j) fft_pass_r4(in, out, tw.data(), NC, Ns, n_fft / (4 * Ns), j);
    std::swap(in, out);
  }
  if (log2nc & 1) {
    for (int j = 0; j < NC / 2; ++j) fft_pass_r2_last(in, out, tw.data(), NC, j);
    std::swap(in, out);
  }
  double maxerr = 0, maxref = 0;
  std::vector<float2> X(NC + 1);
  for (int k = 0; k <= NC; ++k) {
    X[k] = rfft_post(in, tw.data(), NC, k);
    double re = 0, im = 0;
    for (int n = 0; n < n_fft; ++n) {
      re += x[n] * cos(-2.0 * M_PI * k * n / n_fft);
      im += x[n] * sin(-2.0 * M_PI * k * n / n_fft);
    }
    maxerr = fmax(maxerr, hypot(X[k].x - re, X[k].y - im));
    maxref = fmax(maxref, hypot(re, im));
  }
  // inverse
  for (int k = 0; k < NC; ++k) {
    float2 xk = X[k], xn = X[NC - k];
    if (k == 0) { xk.y = 0; xn.y = 0; }
    a[k] = irfft_pre(xk, xn, tw[k]);
  }
  in = a.data(); out = b.data();
  for (int Ns = 1; Ns * 4 <= NC; Ns *= 4) {
    for (int j = 0; j < NC / 4; ++j) fft_pass_r4(in, out, tw.data(), NC, Ns, n_fft / (4 * Ns), j);
    std::swap(in, out);
  }
  if (log2nc & 1) {
    for (int j = 0; j < NC / 2; ++j) fft_pass_r2_last(in, out, tw.data(), NC, j);
    std::swap(in, out);
  }
  double ierr = 0;
  for (int n = 0; n < NC; ++n) {
    ierr = fmax(ierr, fabs(in[n].x / NC - x[2 * n]));
    ierr = fmax(ierr, fabs(-in[n].y / NC - x[2 * n + 1]));
  }
  printf("n_fft %d  fwd max err %.3e (max |X| %.2f)  inverse max err %.3e\n", n_fft, maxerr, maxref, ierr);
  return (maxerr < 2e-5 * maxref + 1e-5 && ierr < 1e-5) ? 0 : 1;
}

// radix 16 x 16 x 4 register-pass FFT (fft1024_*): compare with a double-precision DFT
static int run1024() {
  const int N = 1024;
  std::vector<float2> tw(2048);
  for (int m = 0; m < 2048; ++m) tw[m] = make_float2((float)cos(-2.0 * M_PI * m / 2048), (float)sin(-2.0 * M_PI * m / 2048));
  std::vector<float2> a(FFT1024_BUF), b(FFT1024_BUF), x(N);
  srand(7);
  for (int n = 0; n < N; ++n) x[n] = a[n] = make_float2((float)rand() / RAND_MAX - 0.5f, (float)rand() / RAND_MAX - 0.5f);
  for (int j = 0; j < 64; ++j) fft1024_pass_r16<false, true>(a.data(), b.data(), tw.data(), 1, j);
  for (int j = 0; j < 64; ++j) fft1024_pass_r16<true, true>(b.data(), a.data(), tw.data(), 16, j);
  for (int j = 0; j < 256; ++j) fft1024_pass_r4_last(a.data(), b.data(), tw.data(), j);
  double maxerr = 0, maxref = 0;
  for (int k = 0; k < N; ++k) {
    double re = 0, im = 0;
    for (int n = 0; n < N; ++n) {
      const double c = cos(-2.0 * M_PI * k * n / N), s = sin(-2.0 * M_PI * k * n / N);
      re += x[n].x * c - x[n].y * s;
      im += x[n].x * s + x[n].y * c;
    }
    maxerr = fmax(maxerr, hypot(b[k].x - re, b[k].y - im));
    maxref = fmax(maxref, hypot(re, im));
  }
  printf("fft1024 (16x16x4) max err %.3e (max |X| %.2f)\n", maxerr, maxref);
  return maxerr < 2e-5 * maxref ? 0 : 1;
}

// bin-pair forms used by the staged n_fft = 2048 kernels: rfft_post_pair2 / irfft_pre_pair2 against rfft_post / irfft_pre
static int run_pairs() {
  const int NC = 1024, n_fft = 2048;
  std::vector<float2> tw(n_fft), Z(NC);
  for (int m = 0; m < n_fft; ++m) tw[m] = make_float2((float)cos(-2.0 * M_PI * m / n_fft), (float)sin(-2.0 * M_PI * m / n_fft));
  srand(11);
  for (auto& v : Z) v = make_float2((float)rand() / RAND_MAX - 0.5f, (float)rand() / RAND_MAX - 0.5f);
  double e1 = 0, e2 = 0;
  std::vector<float2> X(NC + 1);
  for (int k = 0; k <= NC; ++k) X[k] = rfft_post(Z.data(), tw.data(), NC, k);
  for (int k = 0; k < NC / 2; ++k) {
    float2 xk, xn;
    rfft_post_pair2(Z[k], Z[(NC - k) & (NC - 1)], tw[k], xk, xn);
    e1 = fmax(e1, hypot(0.5f * xk.x - X[k].x, 0.5f * xk.y - X[k].y));
    e1 = fmax(e1, hypot(0.5f * xn.x - X[NC - k].x, 0.5f * xn.y - X[NC - k].y));
  }
  X[0].y = 0; X[NC].y = 0;
  for (int k = 0; k < NC / 2; ++k) {
    float2 zk2, zn2;
    irfft_pre_pair2(X[k], X[NC - k], tw[k], zk2, zn2);
    const float2 a = irfft_pre(X[k], X[NC - k], tw[k]);
    e2 = fmax(e2, hypot(0.5f * zk2.x - a.x, 0.5f * zk2.y - a.y));
    if (k) {
      const float2 b = irfft_pre(X[NC - k], X[k], tw[NC - k]);
      e2 = fmax(e2, hypot(0.5f * zn2.x - b.x, 0.5f * zn2.y - b.y));
    }
  }
  {
    float2 zk2, zn2;
    irfft_pre_pair2(X[NC / 2], X[NC / 2], tw[NC / 2], zk2, zn2);
    const float2 a = irfft_pre(X[NC / 2], X[NC / 2], tw[NC / 2]);
    e2 = fmax(e2, hypot(0.5f * zk2.x - a.x, 0.5f * zk2.y - a.y));
  }
  printf("pair forms: forward max diff %.3e, inverse max diff %.3e\n", e1, e2);
  return (e1 < 2e-7 && e2 < 2e-7) ? 0 : 1;
}

int main() {
  int bad = run1024() + run_pairs();
  for (int n : {64, 128, 512, 1024, 2048, 4096}) bad += run(n);
  printf(bad ? "FAIL\n" : "OK\n");
  return bad;
}
