// Host-side Fiat-Shamir transcript and proof encoding.
//
// Restates triton-vm/src/proof_stream.rs:19-125 (ProofStream: which items alter the sponge,
// sample_scalars / sample_indices), proof_item.rs:96-147 (variant order, Fiat-Shamir flags) and
// proof.rs:37-88 (Proof = BFieldCodec encoding of the item list; Claim).  The BFieldCodec rules are
// twenty-first 2.0's; they are pinned by the reference's two whole-proof known-answer digests
// (proof.rs:200-226, stark.rs:2433-2460; tests/test_golden.py): struct fields are emitted in reverse declaration order.
//
// Words are kept canonical in the item encodings; the sponge state is Montgomery form.
#pragma once
#include <cstdint>
#include <string>
#include <vector>
#include "field.cuh"
#include "tip5.cuh"

namespace tvm {

static constexpr bool BFIELDCODEC_STRUCT_FIELDS_REVERSED = true;

enum class ItemKind : int {
  MerkleRoot = 0, Log2PaddedHeight = 1, OutOfDomainMainRow = 2, OutOfDomainAuxRow = 3, OutOfDomainQuotientSegments = 4,
  Polynomial = 5, StirOutOfDomainValues = 6, AuthenticationStructure = 7, MasterMainTableRows = 8, MasterAuxTableRows = 9,
  QuotientSegmentsElements = 10, FriCodeword = 11, FriResponse = 12, StirResponse = 13,
};
inline bool item_in_fiat_shamir(ItemKind k) { return (int)k <= 6; }       // proof_item.rs:96-147
inline bool item_payload_static(ItemKind k) { return (int)k <= 4; }

struct Sponge {                       // Tip5::init(): variable-length domain, all-zero state
  u64 s[16] = {0};
  void absorb(const u64 *chunk_mont) {
    for (int i = 0; i < 10; i++) s[i] = chunk_mont[i];
    tip5_permutation_host(s);
  }
  void pad_and_absorb_all(const std::vector<u64> &canon) {
    size_t n = canon.size(), full = n / 10;
    u64 c[10];
    for (size_t b = 0; b < full; b++) {
      for (int i = 0; i < 10; i++) c[i] = to_mont(canon[10 * b + i]);
      absorb(c);
    }
    size_t rem = n - 10 * full;
    for (size_t i = 0; i < 10; i++) c[i] = i < rem ? to_mont(canon[10 * full + i]) : (i == rem ? MONT_ONE : 0);
    absorb(c);
  }
  void squeeze(u64 out_mont[10]) {
    for (int i = 0; i < 10; i++) out_mont[i] = s[i];
    tip5_permutation_host(s);
  }
  // Montgomery-form X-field scalars
  std::vector<xfe> sample_scalars(size_t n) {
    size_t squeezes = (3 * n + 9) / 10;
    std::vector<u64> e;
    for (size_t i = 0; i < squeezes; i++) {
      u64 o[10];
      squeeze(o);
      e.insert(e.end(), o, o + 10);
    }
    std::vector<xfe> out(n);
    for (size_t i = 0; i < n; i++) out[i] = xmake(e[3 * i], e[3 * i + 1], e[3 * i + 2]);
    return out;
  }
  std::vector<uint32_t> sample_indices(uint32_t upper_bound, size_t n) {
    std::vector<uint32_t> out;
    std::vector<u64> buf;  // consumed from the front
    size_t pos = 0;
    while (out.size() != n) {
      if (pos == buf.size()) {
        u64 o[10];
        squeeze(o);
        buf.assign(o, o + 10);
        pos = 0;
      }
      u64 canon = from_mont(buf[pos++]);
      if (canon != P - 1) out.push_back((uint32_t)(canon & 0xFFFFFFFFULL) % upper_bound);
    }
    return out;
  }
};

// A host that keeps `ProofStream` itself ("who keeps the sponge") hands these callbacks to tvm_prove_transcript: the
// library then only PRODUCES proof items and CONSUMES challenges - Fiat-Shamir, the item list and the proof encoding stay with
// the host's own ProofStream / BFieldCodec (proof_stream.rs:40-103).  All words canonical.
struct ExternalTranscript {
  void *user;
  int (*alter_fiat_shamir_state)(void *user, const u64 *words, size_t n);            // ProofStream::alter_fiat_shamir_state_with
  int (*enqueue)(void *user, unsigned variant, const u64 *payload, size_t n);         // ProofStream::enqueue(item): variant index, BFieldCodec payload
  int (*sample_scalars)(void *user, size_t n, u64 *out);                              // [n][3]
  int (*sample_indices)(void *user, unsigned upper_bound, size_t n, unsigned *out);
};
struct TranscriptCallbackError {};

struct ProofStream {
  Sponge sponge;
  std::vector<std::vector<u64>> items;   // full item encodings [variant, (len,) payload...]
  const ExternalTranscript *ext = nullptr;

  std::vector<xfe> sample_scalars(size_t n) {
    if (!ext) return sponge.sample_scalars(n);
    std::vector<u64> w(3 * n);
    if (ext->sample_scalars(ext->user, n, w.data())) throw TranscriptCallbackError{};
    std::vector<xfe> out(n);
    for (size_t i = 0; i < n; i++) out[i] = xmake(to_mont(w[3 * i] % P), to_mont(w[3 * i + 1] % P), to_mont(w[3 * i + 2] % P));
    return out;
  }
  std::vector<uint32_t> sample_indices(uint32_t upper_bound, size_t n) {
    if (!ext) return sponge.sample_indices(upper_bound, n);
    std::vector<uint32_t> out(n);
    if (ext->sample_indices(ext->user, upper_bound, n, out.data())) throw TranscriptCallbackError{};
    for (uint32_t v : out)
      if (v >= upper_bound) throw TranscriptCallbackError{};
    return out;
  }
  void alter_fiat_shamir_state_with(const std::vector<u64> &encoding) {
    if (ext) {
      if (ext->alter_fiat_shamir_state(ext->user, encoding.data(), encoding.size())) throw TranscriptCallbackError{};
      return;
    }
    sponge.pad_and_absorb_all(encoding);
  }
  void enqueue(ItemKind k, const std::vector<u64> &payload) {
    if (ext) {
      if (ext->enqueue(ext->user, (unsigned)(int)k, payload.data(), payload.size())) throw TranscriptCallbackError{};
      return;
    }
    std::vector<u64> e;
    e.reserve(payload.size() + 2);
    e.push_back((u64)(int)k);
    if (!item_payload_static(k)) e.push_back(payload.size());
    e.insert(e.end(), payload.begin(), payload.end());
    if (item_in_fiat_shamir(k)) alter_fiat_shamir_state_with(e);
    items.push_back(std::move(e));
  }
  std::vector<u64> encode() const {     // Proof(Vec<BFE>) (proof_stream.rs:110-119)
    std::vector<u64> body;
    body.push_back(items.size());
    for (auto &it : items) {
      body.push_back(it.size());
      body.insert(body.end(), it.begin(), it.end());
    }
    std::vector<u64> out;
    out.reserve(body.size() + 1);
    out.push_back(body.size());
    out.insert(out.end(), body.begin(), body.end());
    return out;
  }
};

inline void push_xfe_canon(std::vector<u64> &v, xfe x) {
  v.push_back(from_mont(x.c0)); v.push_back(from_mont(x.c1)); v.push_back(from_mont(x.c2));
}

// Claim {program_digest, version, input, output} (proof.rs:68-88)
inline std::vector<u64> encode_claim(const u64 digest[5], uint32_t version, const u64 *in, size_t nin, const u64 *out, size_t nout) {
  std::vector<std::vector<u64>> f(4);
  std::vector<bool> dyn = {false, false, true, true};
  f[0].assign(digest, digest + 5);
  f[1] = {version};
  f[2].push_back(nin); f[2].insert(f[2].end(), in, in + nin);
  f[3].push_back(nout); f[3].insert(f[3].end(), out, out + nout);
  std::vector<u64> e;
  for (int t = 0; t < 4; t++) {
    int i = BFIELDCODEC_STRUCT_FIELDS_REVERSED ? 3 - t : t;
    if (dyn[i]) e.push_back(f[i].size());
    e.insert(e.end(), f[i].begin(), f[i].end());
  }
  return e;
}

// FriResponse {queried_leaves: Vec<XFE>, auth_structure: Vec<Digest>} (fri.rs:99-107); inputs canonical
inline std::vector<u64> encode_fri_response(const std::vector<u64> &leaves_flat, const std::vector<u64> &auth_flat) {
  std::vector<u64> a, b;
  a.push_back(leaves_flat.size() / 3); a.insert(a.end(), leaves_flat.begin(), leaves_flat.end());
  b.push_back(auth_flat.size() / 5); b.insert(b.end(), auth_flat.begin(), auth_flat.end());
  std::vector<u64> e;
  const std::vector<u64> *order[2] = {&a, &b};
  for (int t = 0; t < 2; t++) {
    const std::vector<u64> &f = *order[BFIELDCODEC_STRUCT_FIELDS_REVERSED ? 1 - t : t];
    e.push_back(f.size());
    e.insert(e.end(), f.begin(), f.end());
  }
  return e;
}

}  // namespace tvm
