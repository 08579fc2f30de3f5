// Input side of the path in native code: what vqa/datasets/input_ops_vqa_tf_record_memft.py:17-82 does
// with tf.data on the host -- TFRecord framing, tf.train.Example parsing, parse_fn defaults, sparse_to_dense of the soft
// scores, padded_batch -- for a whole batch per call, straight into the caller's (pinned) batch buffers.
//
// Measured (profiles/r02_input_pipeline.md): the pure-Python mirror (input_ops.create) delivers ~10^4 samples/s, the GPU
// consumes 5 x 10^5; this parser delivers > 10^6 samples/s per host thread. The dense [B, A] soft-score target -- 6.1 MB
// of a 6.2 MB batch -- is NOT built on the host: the parser emits the (row, answer id, score) triples (a few KB) and
// vqa_densify_targets scatters them into the device buffer, so a step uploads ~50 KB instead of 6.2 MB.
//
// Formats restated from their published definitions (no reference code involved; cross-checked against tensorboard's
// TF-team record reader / writer and the protobuf runtime in tests/test_third_party_pin.py):
//   TFRecord : u64 length | u32 masked_crc32c(length) | payload | u32 masked_crc32c(payload), little endian
//   Example  : {1: Features{1: map<string, Feature>}}, Feature = oneof {1: BytesList, 2: FloatList, 3: Int64List},
//              each {1: repeated value}, numeric lists packed or not
#include <cuda_runtime.h>

#include <cstring>

#include "internal.h"
#include "launch.cuh"

namespace vqa {

namespace {

inline uint32_t mask_crc(uint32_t c) { return ((c >> 15) | (c << 17)) + 0xA282EAD8u; }

struct Span {
  const uint8_t* p;
  const uint8_t* end;
};

inline bool read_varint(Span& s, uint64_t* out) {
  uint64_t r = 0;
  for (int shift = 0; shift < 70; shift += 7) {
    if (s.p >= s.end) return false;
    const uint8_t b = *s.p++;
    r |= static_cast<uint64_t>(b & 0x7F) << (shift < 64 ? shift : 63);
    if (!(b & 0x80)) {
      *out = r;
      return true;
    }
  }
  return false;
}

// next field of a message: number, wire type, and for length-delimited fields the payload span
inline bool next_field(Span& s, uint32_t* num, uint32_t* wt, uint64_t* val, Span* sub) {
  uint64_t key;
  if (!read_varint(s, &key)) return false;
  *num = static_cast<uint32_t>(key >> 3);
  *wt = static_cast<uint32_t>(key & 7);
  switch (*wt) {
    case 0: return read_varint(s, val);
    case 1:
      if (s.end - s.p < 8) return false;
      memcpy(val, s.p, 8);
      s.p += 8;
      return true;
    case 2: {
      uint64_t n;
      if (!read_varint(s, &n) || static_cast<uint64_t>(s.end - s.p) < n) return false;
      sub->p = s.p;
      sub->end = s.p + n;
      s.p += n;
      return true;
    }
    case 5: {
      if (s.end - s.p < 4) return false;
      uint32_t v;
      memcpy(&v, s.p, 4);
      *val = v;
      s.p += 4;
      return true;
    }
    default: return false;
  }
}

enum Key { K_QID, K_IMAGE_ID, K_IMAGE_IDX, K_QLIST, K_QLEN, K_AIDS, K_ASCORES, K_OTHER };

inline Key classify(const Span& k) {
  const size_t n = static_cast<size_t>(k.end - k.p);
  auto is = [&](const char* s) { return n == strlen(s) && memcmp(k.p, s, n) == 0; };
  if (is("qid")) return K_QID;
  if (is("image_id")) return K_IMAGE_ID;
  if (is("image_idx")) return K_IMAGE_IDX;
  if (is("q_intseq/list")) return K_QLIST;
  if (is("q_intseq/len")) return K_QLEN;
  if (is("answers/ids")) return K_AIDS;
  if (is("answers/scores")) return K_ASCORES;
  return K_OTHER;
}

// Int64List payload -> up to cap values; returns the count or -1 (malformed) / -2 (more than cap)
inline int read_int64_list(Span list, long long* out, int cap) {
  int n = 0;
  uint32_t num, wt;
  uint64_t val;
  Span sub{};
  while (list.p < list.end) {
    if (!next_field(list, &num, &wt, &val, &sub)) return -1;
    if (num != 1) continue;
    if (wt == 2) {
      while (sub.p < sub.end) {
        uint64_t v;
        if (!read_varint(sub, &v)) return -1;
        if (n >= cap) return -2;
        out[n++] = static_cast<long long>(v);
      }
    } else if (wt == 0) {
      if (n >= cap) return -2;
      out[n++] = static_cast<long long>(val);
    }
  }
  return n;
}

inline int read_float_list(Span list, float* out, int cap) {
  int n = 0;
  uint32_t num, wt;
  uint64_t val;
  Span sub{};
  while (list.p < list.end) {
    if (!next_field(list, &num, &wt, &val, &sub)) return -1;
    if (num != 1) continue;
    if (wt == 2) {
      const size_t cnt = static_cast<size_t>(sub.end - sub.p) / 4;
      if (static_cast<size_t>(sub.end - sub.p) % 4) return -1;
      if (n + static_cast<int>(cnt) > cap) return -2;
      memcpy(out + n, sub.p, cnt * 4);
      n += static_cast<int>(cnt);
    } else if (wt == 5) {
      if (n >= cap) return -2;
      const uint32_t bits = static_cast<uint32_t>(val);
      memcpy(out + n, &bits, 4);
      ++n;
    }
  }
  return n;
}

constexpr int kMaxAnswers = 64;   // per sample (VQA v2: <= 10 human answers; generator_tf_record_memft_genome.py:122-132)
constexpr int kMaxTokens = 256;

// soft-score targets: target[row, id] = score for every triple, zero elsewhere (tf.sparse_to_dense, default 0)
__global__ void densify_kernel(const int* __restrict__ rows, const int* __restrict__ ids, const float* __restrict__ scores,
                               int n, int A, float* __restrict__ out) {
  pdl_sync();
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
    out[static_cast<long long>(rows[i]) * A + ids[i]] = scores[i];
}

}  // namespace

}  // namespace vqa

using namespace vqa;

extern "C" {

VQA_API VqaStatus vqa_tfrecord_index_host(const uint8_t* file_host, uint64_t size, int32_t verify_crc, uint64_t* offsets,
                                          uint64_t* lengths, int64_t capacity, int64_t* count) {
  if (!file_host || !count || (capacity > 0 && (!offsets || !lengths)))
    return set_error(VQA_ERR_BAD_ARG, "vqa_tfrecord_index_host: null argument");
  uint64_t pos = 0;
  int64_t n = 0;
  while (pos < size) {
    if (size - pos < 12) return set_error(VQA_ERR_BAD_SHAPE, "TFRecord: truncated record header at byte %llu", static_cast<unsigned long long>(pos));
    uint64_t len;
    uint32_t lcrc;
    memcpy(&len, file_host + pos, 8);
    memcpy(&lcrc, file_host + pos + 8, 4);
    if (verify_crc && mask_crc(vqa_crc32c(file_host + pos, 8)) != lcrc)
      return set_error(VQA_ERR_BAD_SHAPE, "TFRecord: corrupted record length at byte %llu", static_cast<unsigned long long>(pos));
    if (size - pos - 12 < len + 4) return set_error(VQA_ERR_BAD_SHAPE, "TFRecord: truncated record at byte %llu", static_cast<unsigned long long>(pos));
    if (verify_crc) {
      uint32_t dcrc;
      memcpy(&dcrc, file_host + pos + 12 + len, 4);
      if (mask_crc(vqa_crc32c(file_host + pos + 12, len)) != dcrc)
        return set_error(VQA_ERR_BAD_SHAPE, "TFRecord: corrupted record payload at byte %llu", static_cast<unsigned long long>(pos));
    }
    if (n < capacity) {
      offsets[n] = pos + 12;
      lengths[n] = len;
    }
    ++n;
    pos += 12 + len + 4;
  }
  *count = n;   // (> capacity: call again with larger arrays)
  return VQA_OK;
}

VQA_API VqaStatus vqa_parse_examples_host(const uint8_t* const* records, const uint64_t* lengths, int32_t n,
                                          int32_t num_answers, int32_t t_cap, int64_t* id, int64_t* image_idx,
                                          int32_t* q_intseq, int32_t* q_intseq_len, int32_t* t_longest,
                                          int32_t* ans_row, int32_t* ans_id, float* ans_score, int32_t ans_cap,
                                          int32_t* ans_count, uint32_t* image_id_off, uint32_t* image_id_len) {
  if (!records || !lengths || n < 0 || !id || !image_idx || !q_intseq || !q_intseq_len || !t_longest || !ans_row ||
      !ans_id || !ans_score || !ans_count || t_cap <= 0 || t_cap > kMaxTokens)
    return set_error(VQA_ERR_BAD_ARG, "vqa_parse_examples_host: null / bad argument");
  int longest = 0, na = 0;
  for (int i = 0; i < n; ++i) {
    Span ex{records[i], records[i] + lengths[i]};
    long long qid = -1, iidx = -1, qlen = -1;          // parse_fn defaults: qid -1, image_idx -1; q_intseq/len is required
    bool have_qlen = false;
    long long toks[kMaxTokens];
    int ntok = 0;
    long long aid[kMaxAnswers];
    float asc[kMaxAnswers];
    int nid = 0, nsc = 0;
    uint32_t iid_off = 0, iid_len = 0;
    uint32_t num, wt;
    uint64_t val;
    Span feats{}, entry{}, sub{};
    while (ex.p < ex.end) {
      if (!next_field(ex, &num, &wt, &val, &feats)) return set_error(VQA_ERR_BAD_SHAPE, "Example %d: malformed message", i);
      if (num != 1 || wt != 2) continue;
      while (feats.p < feats.end) {
        if (!next_field(feats, &num, &wt, &val, &entry)) return set_error(VQA_ERR_BAD_SHAPE, "Example %d: malformed Features", i);
        if (num != 1 || wt != 2) continue;
        Span key{nullptr, nullptr}, feat{nullptr, nullptr};
        while (entry.p < entry.end) {
          if (!next_field(entry, &num, &wt, &val, &sub)) return set_error(VQA_ERR_BAD_SHAPE, "Example %d: malformed map entry", i);
          if (wt != 2) continue;
          if (num == 1) key = sub;
          else if (num == 2) feat = sub;
        }
        if (!key.p) continue;
        const Key k = classify(key);
        if (k == K_OTHER || !feat.p) continue;
        // the Feature's oneof: 1 bytes_list, 2 float_list, 3 int64_list
        Span f = feat, list{};
        while (f.p < f.end) {
          if (!next_field(f, &num, &wt, &val, &list)) return set_error(VQA_ERR_BAD_SHAPE, "Example %d: malformed Feature", i);
          if (wt != 2) continue;
          if (num == 3 && (k == K_QID || k == K_IMAGE_IDX || k == K_QLEN)) {
            long long v1[2];
            const int c = read_int64_list(list, v1, 1);
            if (c == -1) return set_error(VQA_ERR_BAD_SHAPE, "Example %d: malformed Int64List", i);
            if (c >= 1 || c == -2) {
              if (k == K_QID) qid = v1[0];
              else if (k == K_IMAGE_IDX) iidx = v1[0];
              else { qlen = v1[0]; have_qlen = true; }
            }
          } else if (num == 3 && k == K_QLIST) {
            ntok = read_int64_list(list, toks, t_cap);
            if (ntok == -1) return set_error(VQA_ERR_BAD_SHAPE, "Example %d: malformed q_intseq/list", i);
            if (ntok == -2) return set_error(VQA_ERR_BAD_SHAPE, "Example %d: question longer than the %d-token row", i, t_cap);
          } else if (num == 3 && k == K_AIDS) {
            nid = read_int64_list(list, aid, kMaxAnswers);
            if (nid < 0) return set_error(VQA_ERR_BAD_SHAPE, "Example %d: malformed / oversized answers/ids", i);
          } else if (num == 2 && k == K_ASCORES) {
            nsc = read_float_list(list, asc, kMaxAnswers);
            if (nsc < 0) return set_error(VQA_ERR_BAD_SHAPE, "Example %d: malformed / oversized answers/scores", i);
          } else if (num == 1 && k == K_IMAGE_ID) {
            Span bl = list, one{};
            while (bl.p < bl.end) {
              if (!next_field(bl, &num, &wt, &val, &one)) return set_error(VQA_ERR_BAD_SHAPE, "Example %d: malformed BytesList", i);
              if (num == 1 && wt == 2) {
                iid_off = static_cast<uint32_t>(one.p - records[i]);
                iid_len = static_cast<uint32_t>(one.end - one.p);
                break;
              }
            }
          }
        }
      }
    }
    if (!have_qlen) return set_error(VQA_ERR_BAD_SHAPE, "Example %d: feature 'q_intseq/len' is required", i);
    if (nid != nsc) return set_error(VQA_ERR_BAD_SHAPE, "Example %d: answers/ids and answers/scores differ in length", i);
    id[i] = qid;
    image_idx[i] = iidx;
    q_intseq_len[i] = static_cast<int32_t>(qlen);
    int32_t* row = q_intseq + static_cast<long long>(i) * t_cap;
    for (int t = 0; t < ntok; ++t) row[t] = static_cast<int32_t>(toks[t]);
    for (int t = ntok; t < t_cap; ++t) row[t] = 0;      // padded_batch pads with 0
    if (ntok > longest) longest = ntok;
    if (image_id_off) { image_id_off[i] = iid_off; image_id_len[i] = iid_len; }
    for (int a = 0; a < nid; ++a) {
      if (aid[a] < 0 || aid[a] >= num_answers) return set_error(VQA_ERR_BAD_SHAPE, "Example %d: answer id %lld out of range", i, aid[a]);
      // target[ids] = scores: a repeated id keeps its LAST score -- drop the earlier triples so the device scatter has no race
      bool later = false;
      for (int a2 = a + 1; a2 < nid; ++a2) later |= aid[a2] == aid[a];
      if (later) continue;
      if (na >= ans_cap) return set_error(VQA_ERR_BAD_SHAPE, "vqa_parse_examples_host: more than %d answer triples in the batch", ans_cap);
      ans_row[na] = i;
      ans_id[na] = static_cast<int32_t>(aid[a]);
      ans_score[na] = asc[a];
      ++na;
    }
  }
  *t_longest = longest;
  *ans_count = na;
  return VQA_OK;
}

VQA_API VqaStatus vqa_densify_targets(const int32_t* rows, const int32_t* ids, const float* scores, int32_t n, int32_t batch,
                                      int32_t num_answers, float* target, void* stream) {
  if (!target || batch < 0 || num_answers <= 0 || n < 0 || (n > 0 && (!rows || !ids || !scores)))
    return set_error(VQA_ERR_BAD_ARG, "vqa_densify_targets: bad argument");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (batch == 0) return VQA_OK;
  VQA_CUDA_CHECK(cudaMemsetAsync(target, 0, sizeof(float) * static_cast<size_t>(batch) * num_answers, s));
  if (n > 0) {
    launch_pdl(densify_kernel, dim3((n + 255) / 256), dim3(256), 0, s, rows, ids, scores, n, num_answers, target);
    VQA_LAUNCH_CHECK("densify_targets");
  }
  return VQA_OK;
}

}  // extern "C"
