// On-disk container for encodings (a CRS range, a whole proving key arena, or a proof) -- SURVEY.md 8(f) rank 4.
// The reference declares proving-key / proof stream operators (zk_proof_systems/r1cs_ppzksnark.hpp:43-47,142-146,
// groth16.hpp:23-27) and never defines them, so there is no format to be compatible with; this one is the HBM layout
// written down as it is:
//   8 bytes  magic "RSGKEY01"
//   u64      kind (1 = CRS / proving-key range, 2 = proof), N_R, L_R, N_E, L_E, n_elems
//   u64[L_R] ring primes q_j;  u64[L_E] encoding primes Q_l
//   payload  n_elems encodings, each [L_R][2][L_E][N_E] u64 words (canonical residues, NTT form, first level)
//   u64      checksum of the payload (word-wise FNV-1a-64), u64 end mark 0x444E454B47535221 ("!RSGKEND")
// All integers little endian.  A reader rejects: wrong magic / end mark, a parameter or prime that differs from the
// context's, a word >= its prime, a checksum mismatch, a short file.  Host-only entry points (rsg_enc_file_*) need no
// GPU; rsg_crs_save / rsg_crs_load stream between the file and HBM through a pinned staging buffer.
#include <cstdio>

namespace {
constexpr uint64_t kFileEnd = 0x444E454B47535221ull;
constexpr size_t kHeadWords = 6;
inline uint64_t fnv_words(uint64_t h, const uint64_t *w, size_t n) {
  for (size_t i = 0; i < n; i++) h = (h ^ w[i]) * 0x100000001b3ull;
  return h;
}
constexpr uint64_t kFnvInit = 0xcbf29ce484222325ull;
struct FileCloser {
  FILE *f;
  ~FileCloser() { if (f) fclose(f); }
  // writers: a buffered write can fail as late as here (ENOSPC); report it instead of returning RSG_OK
  int finish() {
    FILE *g = f;
    f = nullptr;
    if (!g) return RSG_OK;
    const bool bad = fflush(g) != 0;
    return (fclose(g) != 0 || bad) ? fail(RSG_ERR_ARG, "write failed (flush / close)") : RSG_OK;
  }
};
// header of an open file -> info[0..5] = kind, N_R, L_R, N_E, L_E, n_elems; primes appended to q / Q
int read_header(FILE *f, uint64_t *info, std::vector<uint64_t> &q, std::vector<uint64_t> &Q) {
  char magic[8];
  if (fread(magic, 1, 8, f) != 8 || memcmp(magic, "RSGKEY01", 8)) return fail(RSG_ERR_ARG, "not an RSGKEY01 file");
  if (fread(info, 8, kHeadWords, f) != kHeadWords) return fail(RSG_ERR_ARG, "short header");
  if (info[2] == 0 || info[2] > MAX_LR || info[4] == 0 || info[4] > MAX_LE) return fail(RSG_ERR_ARG, "bad limb counts in header");
  q.resize(info[2]);
  Q.resize(info[4]);
  if (fread(q.data(), 8, q.size(), f) != q.size() || fread(Q.data(), 8, Q.size(), f) != Q.size()) return fail(RSG_ERR_ARG, "short header");
  return RSG_OK;
}
int write_header(FILE *f, uint64_t kind, uint64_t N_R, uint64_t L_R, const uint64_t *q, uint64_t N_E, uint64_t L_E, const uint64_t *Q,
                 uint64_t n_elems) {
  const uint64_t head[kHeadWords] = {kind, N_R, L_R, N_E, L_E, n_elems};
  if (fwrite("RSGKEY01", 1, 8, f) != 8 || fwrite(head, 8, kHeadWords, f) != kHeadWords || fwrite(q, 8, L_R, f) != L_R ||
      fwrite(Q, 8, L_E, f) != L_E)
    return fail(RSG_ERR_ARG, "write failed");
  return RSG_OK;
}
// every word of `count` encodings below its prime
bool words_in_range(const uint64_t *w, size_t count, size_t L_R, size_t L_E, size_t N_E, const uint64_t *Q) {
  for (size_t e = 0; e < count * L_R * 2; e++)
    for (size_t l = 0; l < L_E; l++) {
      const uint64_t *row = w + (e * L_E + l) * N_E, p = Q[l];
      for (size_t i = 0; i < N_E; i++)
        if (row[i] >= p) return false;
    }
  return true;
}
}  // namespace

extern "C" int rsg_enc_file_info(const char *path, uint64_t *info /* 6 */, uint64_t *q /* MAX 8, nullable */, uint64_t *Q /* MAX 16, nullable */) {
  if (!path || !info) return fail(RSG_ERR_ARG, "null argument");
  FileCloser fc{fopen(path, "rb")};
  if (!fc.f) return fail(RSG_ERR_ARG, "cannot open file");
  std::vector<uint64_t> vq, vQ;
  int rc = read_header(fc.f, info, vq, vQ);
  if (rc) return rc;
  if (q) memcpy(q, vq.data(), vq.size() * 8);
  if (Q) memcpy(Q, vQ.data(), vQ.size() * 8);
  return RSG_OK;
}

extern "C" int rsg_enc_file_write(const char *path, uint64_t kind, size_t N_R, size_t L_R, const uint64_t *q, size_t N_E, size_t L_E,
                                  const uint64_t *Q, size_t n_elems, const uint64_t *h_words) {
  if (!path || !q || !Q || (!h_words && n_elems)) return fail(RSG_ERR_ARG, "null argument");
  if (!L_R || L_R > MAX_LR || !L_E || L_E > MAX_LE) return fail(RSG_ERR_ARG, "limb counts");
  const size_t words = n_elems * L_R * 2 * L_E * N_E;
  if (!words_in_range(h_words, n_elems, L_R, L_E, N_E, Q)) return fail(RSG_ERR_ARG, "a word is not a canonical residue");
  FileCloser fc{fopen(path, "wb")};
  if (!fc.f) return fail(RSG_ERR_ARG, "cannot create file");
  int rc = write_header(fc.f, kind, N_R, L_R, q, N_E, L_E, Q, n_elems);
  if (rc) return rc;
  const uint64_t tail[2] = {fnv_words(kFnvInit, h_words, words), kFileEnd};
  if (fwrite(h_words, 8, words, fc.f) != words || fwrite(tail, 8, 2, fc.f) != 2) return fail(RSG_ERR_ARG, "write failed");
  if (int frc = fc.finish()) return frc;
  return RSG_OK;
}

extern "C" int rsg_enc_file_read(const char *path, size_t N_R, size_t L_R, const uint64_t *q, size_t N_E, size_t L_E, const uint64_t *Q,
                                 size_t cap_elems, uint64_t *h_words, size_t *n_elems, uint64_t *kind) {
  if (!path || !q || !Q || !h_words || !n_elems) return fail(RSG_ERR_ARG, "null argument");
  FileCloser fc{fopen(path, "rb")};
  if (!fc.f) return fail(RSG_ERR_ARG, "cannot open file");
  uint64_t info[kHeadWords];
  std::vector<uint64_t> vq, vQ;
  int rc = read_header(fc.f, info, vq, vQ);
  if (rc) return rc;
  if (info[1] != N_R || info[2] != L_R || info[3] != N_E || info[4] != L_E || memcmp(vq.data(), q, L_R * 8) || memcmp(vQ.data(), Q, L_E * 8))
    return fail(RSG_ERR_ARG, "file was written for other parameters");
  if (info[5] > cap_elems) return fail(RSG_ERR_ARG, "buffer too small for the file's encodings");
  const size_t words = info[5] * L_R * 2 * L_E * N_E;
  uint64_t tail[2];
  if (fread(h_words, 8, words, fc.f) != words || fread(tail, 8, 2, fc.f) != 2) return fail(RSG_ERR_ARG, "short file");
  if (tail[1] != kFileEnd) return fail(RSG_ERR_ARG, "end mark missing");
  if (tail[0] != fnv_words(kFnvInit, h_words, words)) return fail(RSG_ERR_ARG, "checksum mismatch");
  if (!words_in_range(h_words, info[5], L_R, L_E, N_E, Q)) return fail(RSG_ERR_ARG, "a word is not a canonical residue");
  *n_elems = info[5];
  if (kind) *kind = info[0];
  return RSG_OK;
}

// HBM <-> file, `chunk` encodings at a time through pinned memory
extern "C" int rsg_crs_save(const rsg_crs *r, size_t first, size_t count, const char *path) {
  if (!r || !path || first + count > r->n) return fail(RSG_ERR_ARG, "CRS range");
  rsg_context *c = r->ctx;
  std::lock_guard<std::mutex> g(c->mu);
  CUDA_TRY(cudaSetDevice(c->device));
  FileCloser fc{fopen(path, "wb")};
  if (!fc.f) return fail(RSG_ERR_ARG, "cannot create file");
  int rc = write_header(fc.f, 1, c->N_R, c->L_R, c->q.data(), c->N_E, c->L_E, c->Q.data(), count);
  if (rc) return rc;
  const size_t ew = c->enc_words(), chunk = std::max<size_t>(1, ((size_t)64 << 20) / (ew * 8));
  uint64_t *stage = nullptr;
  CUDA_TRY(cudaMallocHost((void **)&stage, chunk * ew * 8));
  uint64_t h = kFnvInit;
  rc = RSG_OK;
  for (size_t i = 0; i < count && rc == RSG_OK; i += chunk) {
    const size_t k = std::min(chunk, count - i);
    if (cudaMemcpyAsync(stage, r->d + (first + i) * ew, k * ew * 8, cudaMemcpyDeviceToHost, c->stream) != cudaSuccess ||
        cudaStreamSynchronize(c->stream) != cudaSuccess)
      rc = fail(RSG_ERR_CUDA, "device to host copy failed");
    else if (fwrite(stage, 8, k * ew, fc.f) != k * ew)
      rc = fail(RSG_ERR_ARG, "write failed");
    else
      h = fnv_words(h, stage, k * ew);
  }
  cudaFreeHost(stage);
  if (rc) return rc;
  const uint64_t tail[2] = {h, kFileEnd};
  if (fwrite(tail, 8, 2, fc.f) != 2) return fail(RSG_ERR_ARG, "write failed");
  return fc.finish();
}

extern "C" int rsg_crs_load(rsg_crs *r, size_t first, const char *path, size_t *count_out) {
  if (!r || !path) return fail(RSG_ERR_ARG, "null argument");
  rsg_context *c = r->ctx;
  std::lock_guard<std::mutex> g(c->mu);
  CUDA_TRY(cudaSetDevice(c->device));
  FileCloser fc{fopen(path, "rb")};
  if (!fc.f) return fail(RSG_ERR_ARG, "cannot open file");
  uint64_t info[kHeadWords];
  std::vector<uint64_t> vq, vQ;
  int rc = read_header(fc.f, info, vq, vQ);
  if (rc) return rc;
  if (info[1] != c->N_R || info[2] != c->L_R || info[3] != c->N_E || info[4] != c->L_E || vq != c->q || vQ != c->Q)
    return fail(RSG_ERR_ARG, "file was written for other parameters");
  const size_t count = info[5];
  if (first + count > r->n) return fail(RSG_ERR_ARG, "arena too small for the file's encodings");
  const size_t ew = c->enc_words(), chunk = std::max<size_t>(1, ((size_t)64 << 20) / (ew * 8));
  uint64_t *stage = nullptr;
  CUDA_TRY(cudaMallocHost((void **)&stage, chunk * ew * 8));
  // pass 0 verifies the whole payload (ranges, checksum, end mark) WITHOUT touching the arena; pass 1 copies.  A corrupt or
  // truncated file therefore leaves the proving key as it was.
  const long payload = ftell(fc.f);
  for (int pass = 0; pass < 2 && rc == RSG_OK; pass++) {
    if (fseek(fc.f, payload, SEEK_SET) != 0) { rc = fail(RSG_ERR_ARG, "seek failed"); break; }
    uint64_t h = kFnvInit;
    for (size_t i = 0; i < count && rc == RSG_OK; i += chunk) {
      const size_t k = std::min(chunk, count - i);
      if (fread(stage, 8, k * ew, fc.f) != k * ew)
        rc = fail(RSG_ERR_ARG, "short file");
      else if (pass == 0 && !words_in_range(stage, k, c->L_R, c->L_E, c->N_E, c->Q.data()))
        rc = fail(RSG_ERR_ARG, "a word is not a canonical residue");
      else if (pass == 1 && (cudaMemcpyAsync(r->d + (first + i) * ew, stage, k * ew * 8, cudaMemcpyHostToDevice, c->stream) != cudaSuccess ||
                             cudaStreamSynchronize(c->stream) != cudaSuccess))
        rc = fail(RSG_ERR_CUDA, "host to device copy failed");
      else
        h = fnv_words(h, stage, k * ew);
    }
    if (rc) break;
    uint64_t tail[2];
    if (fread(tail, 8, 2, fc.f) != 2 || tail[1] != kFileEnd) rc = fail(RSG_ERR_ARG, "end mark missing");
    else if (tail[0] != h) rc = fail(RSG_ERR_ARG, pass == 0 ? "checksum mismatch (arena untouched)" : "file changed between the two passes");
  }
  cudaFreeHost(stage);
  if (rc) return rc;
  if (count_out) *count_out = count;
  return RSG_OK;
}
