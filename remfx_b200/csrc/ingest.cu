// Batch ingest: the step in front of the hot path.
//
// The reference feeds its networks from pre-rendered chunk directories: `EffectDataset.__getitem__`
// (remfx/datasets.py:461-468) does two `torchaudio.load` calls per item (`input.wav`, `target.wav`: mono RIFF/WAVE files
// written by `torchaudio.save`, 32-bit float at the reference's settings, remfx/datasets.py:197-198,447-448) and the
// DataLoader (8 worker processes, cfg/config.yaml:105-107) collates (1, T) tensors into a (B, 1, T) batch that is then
// copied to the device.  At B200 speeds (one 32 x 262144 batch every ~1.7 ms) that path is the limiter, so here a batch is
// decoded by a small thread pool straight into ONE pinned (B, 1, T) staging buffer that `rfx_umx_pipe_push` consumes with
// x_on_host = 1: no per-item tensors, no collate copy, no pageable memory.
//
// Host code only (this file holds no kernel); plain C ABI like the rest of the library.
#include "common.cuh"
#include "../../include/remfx_b200.h"

#include <atomic>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

namespace {

struct WavInfo {
  int format = 0;  // 1 = integer PCM, 3 = IEEE float
  int channels = 0, sample_rate = 0, bits = 0;
  long long frames = 0;
  long long data_off = 0;
};

uint32_t rd32(const unsigned char* p) { return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24); }
uint16_t rd16(const unsigned char* p) { return (uint16_t)(p[0] | (p[1] << 8)); }

// Walks the RIFF chunks: "fmt " (PCM / IEEE float / WAVE_FORMAT_EXTENSIBLE) and "data"; every other chunk (LIST, fact, PEAK, ...)
// is skipped.  Returns an empty string on success, else the reason.
std::string parse_header(FILE* f, WavInfo* w) {
  unsigned char h[12];
  if (fread(h, 1, 12, f) != 12 || memcmp(h, "RIFF", 4) != 0 || memcmp(h + 8, "WAVE", 4) != 0) return "not a RIFF/WAVE file";
  bool have_fmt = false;
  for (;;) {
    unsigned char ch[8];
    if (fread(ch, 1, 8, f) != 8) return have_fmt ? "no data chunk" : "no fmt chunk";
    const uint32_t size = rd32(ch + 4);
    if (memcmp(ch, "fmt ", 4) == 0) {
      unsigned char b[40] = {0};
      const uint32_t n = size < 40 ? size : 40;
      if (size < 16 || fread(b, 1, n, f) != n) return "truncated fmt chunk";
      w->format = rd16(b);
      w->channels = rd16(b + 2);
      w->sample_rate = (int)rd32(b + 4);
      w->bits = rd16(b + 14);
      if (w->format == 0xFFFE) {  // WAVE_FORMAT_EXTENSIBLE: the first two bytes of the sub-format GUID are the real tag
        if (size < 40) return "truncated extensible fmt chunk";
        w->format = rd16(b + 24);
      }
      if (fseek(f, (long)(size - n) + (long)(size & 1), SEEK_CUR) != 0) return "seek failed";
      have_fmt = true;
    } else if (memcmp(ch, "data", 4) == 0) {
      if (!have_fmt) return "data chunk before fmt chunk";
      w->data_off = ftell(f);
      const int bytes = w->bits / 8;
      if (w->channels < 1 || bytes < 1) return "bad channel count / sample size";
      w->frames = (long long)size / ((long long)bytes * w->channels);
      return "";
    } else {
      if (fseek(f, (long)size + (long)(size & 1), SEEK_CUR) != 0) return "seek failed";
    }
  }
}

// Decodes `frames` mono samples to float, the way torchaudio.load(normalize=True) does: IEEE float as is, integer PCM scaled by
// 2^-(bits-1) (8-bit WAV is unsigned with a 128 offset).
std::string decode(FILE* f, const WavInfo& w, float* dst, long long frames) {
  if (w.channels != 1) return "expected a mono file (the reference renders mono chunks, remfx/datasets.py:441-442), got " + std::to_string(w.channels) + " channels";
  if (fseek(f, (long)w.data_off, SEEK_SET) != 0) return "seek failed";
  if (w.format == 3 && w.bits == 32) {
    if ((long long)fread(dst, 4, (size_t)frames, f) != frames) return "truncated data chunk";
    return "";
  }
  if (w.format == 3 && w.bits == 64) {
    std::vector<double> tmp((size_t)frames);
    if ((long long)fread(tmp.data(), 8, (size_t)frames, f) != frames) return "truncated data chunk";
    for (long long i = 0; i < frames; ++i) dst[i] = (float)tmp[(size_t)i];
    return "";
  }
  if (w.format == 1 && (w.bits == 8 || w.bits == 16 || w.bits == 24 || w.bits == 32)) {
    const int bytes = w.bits / 8;
    std::vector<unsigned char> tmp((size_t)frames * bytes);
    if ((long long)fread(tmp.data(), (size_t)bytes, (size_t)frames, f) != frames) return "truncated data chunk";
    const unsigned char* p = tmp.data();
    for (long long i = 0; i < frames; ++i, p += bytes) {
      if (bytes == 1) dst[i] = ((int)p[0] - 128) * (1.0f / 128.0f);
      else if (bytes == 2) dst[i] = (float)(int16_t)rd16(p) * (1.0f / 32768.0f);
      else if (bytes == 3) dst[i] = (float)(((int32_t)((uint32_t)p[0] << 8 | (uint32_t)p[1] << 16 | (uint32_t)p[2] << 24)) >> 8) * (1.0f / 8388608.0f);
      else dst[i] = (float)((double)(int32_t)rd32(p) * (1.0 / 2147483648.0));
    }
    return "";
  }
  return "unsupported sample format (tag " + std::to_string(w.format) + ", " + std::to_string(w.bits) + " bits)";
}

std::string read_one(const char* path, float* dst, long long T, long long* frames_out, int* sr_out) {
  FILE* f = fopen(path, "rb");
  if (!f) return std::string("cannot open ") + path;
  WavInfo w;
  std::string err = parse_header(f, &w);
  if (err.empty()) {
    const long long n = w.frames < T ? w.frames : T;
    err = decode(f, w, dst, n);
    if (err.empty()) {
      for (long long i = n; i < T; ++i) dst[i] = 0.0f;  // short file: zero-padded row; `frames` tells the caller (the reference's chunk files all hold chunk_size samples)
      if (frames_out) *frames_out = w.frames;
      if (sr_out) *sr_out = w.sample_rate;
    }
  }
  fclose(f);
  return err.empty() ? err : std::string(path) + ": " + err;
}

}  // namespace

extern "C" {

int rfx_wav_info(const char* path, int* sample_rate, int* channels, long long* frames, int* format_tag, int* bits) {
  RFX_REQUIRE(path, "null path");
  FILE* f = fopen(path, "rb");
  if (!f) { rfx::set_error(std::string("cannot open ") + path); return 2; }
  WavInfo w;
  const std::string err = parse_header(f, &w);
  fclose(f);
  if (!err.empty()) { rfx::set_error(std::string(path) + ": " + err); return 2; }
  if (sample_rate) *sample_rate = w.sample_rate;
  if (channels) *channels = w.channels;
  if (frames) *frames = w.frames;
  if (format_tag) *format_tag = w.format;
  if (bits) *bits = w.bits;
  return 0;
}

int rfx_ingest_wav_batch(const char* const* paths, int n, float* dst_host, long long T, int n_threads, long long* frames, int* sample_rates) {
  RFX_REQUIRE(paths && dst_host && n > 0 && T > 0, "bad argument");
  const int nt = n_threads < 1 ? 1 : (n_threads > n ? n : n_threads);
  std::vector<std::string> errs((size_t)n);
  std::atomic<int> next{0};
  auto worker = [&]() {
    for (int i = next.fetch_add(1); i < n; i = next.fetch_add(1))
      errs[(size_t)i] = read_one(paths[i], dst_host + (size_t)i * (size_t)T, T, frames ? frames + i : nullptr, sample_rates ? sample_rates + i : nullptr);
  };
  std::vector<std::thread> pool;
  for (int t = 1; t < nt; ++t) pool.emplace_back(worker);
  worker();
  for (auto& th : pool) th.join();
  for (int i = 0; i < n; ++i)
    if (!errs[(size_t)i].empty()) { rfx::set_error(errs[(size_t)i]); return 2; }
  return 0;
}

}  // extern "C"
