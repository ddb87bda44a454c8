// transfer_kernels.cuh -- field-masked transfers between the packed host records (simulation.agents.array, 228 / 316 B per
// agent, every f8 at an offset = 4 mod 8) and the device planes, done by kernels that read / write PINNED HOST MEMORY
// directly (zero-copy over PCIe).
//
// Measured on this pool's B200 hosts (profiles/pcie_probe_r2.txt, 1 M three-circle records): a whole-record DMA moves 316 B
// per agent at 55.6 / 52.6 GB/s (5.7 / 6.0 ms); strided DMA (cudaMemcpy2D over the three mutable spans) and zero-copy
// kernels move the 152 mutable bytes at 18 - 26 GB/s (5.8 - 8.3 ms) -- no faster than the whole record -- but their time
// scales with the BYTES SELECTED, so a node that wrote 16 - 104 B per agent (every node of the reference tree) moves them
// in 0.8 - 5 ms instead of 6.  Rule used by the C ABI: masks up to TRANSFER_ZERO_COPY_MAX bytes go through these kernels
// when the host array is pinned / registered, everything else through the whole-record DMA.
//
// Thread mapping as in the probe: consecutive threads handle consecutive 32-bit words of ONE record's selected fields, so a
// warp touches a few contiguous host segments (PCIe payloads of 32 - 128 B) instead of 32 scattered words.
#pragma once
#include "kernels.cuh"

constexpr int TRANSFER_ZERO_COPY_MAX = 120;     // bytes per agent
constexpr int MAX_FIELD_WORDS = 2 * N_FIELDS_THREE;

struct WordMap {                 // the selected fields of one record, as 32-bit words in record order
    int n_words;
    short word[MAX_FIELD_WORDS];     // word offset inside the record
    short plane[MAX_FIELD_WORDS];    // plane of the double the word belongs to
    unsigned char hi[MAX_FIELD_WORDS];   // 1: upper half of the double
};

// device planes -> host records: thread g = (slot t, word w)
__global__ void k_fields_to_host(Soa s, int n, uint32_t *__restrict__ host, int item_words, const WordMap m) {
    const long long g = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    const int t = (int)(g / m.n_words), w = (int)(g % m.n_words);
    if (t >= n) return;
    const int rec = s.id[t];
    if (rec < 0) return;
    const double v = s(m.plane[w], t);
    host[(size_t)rec * item_words + m.word[w]] = m.hi[w] ? (uint32_t)__double2hiint(v) : (uint32_t)__double2loint(v);
}

// host records -> device planes: thread g = (record, double d of the selected fields); two 4-byte host reads per double
__global__ void k_fields_from_host(const uint32_t *__restrict__ host, int item_words, Soa s, int n, const int *__restrict__ slot_of_id,
                                   const WordMap m) {
    const int nd = m.n_words / 2;
    const long long g = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    const int rec = (int)(g / nd), d = (int)(g % nd);
    if (rec >= n) return;
    const int t = slot_of_id ? slot_of_id[rec] : rec;
    const uint32_t *p = host + (size_t)rec * item_words + m.word[2 * d];
    s(m.plane[2 * d], t) = __hiloint2double((int)p[1], (int)p[0]);
}

__global__ void k_slot_of_id(Soa s, int n, int *__restrict__ slot_of_id) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < n && s.id[t] >= 0) slot_of_id[s.id[t]] = t;
}

// whole-record snapshot (cdb_snapshot_begin): `aos` already holds a copy of the uploaded image; every mutable field, the
// States.target and -- where the device owns them -- active / is_follower / index_leader are overwritten with current values
template <int MODEL>
__global__ void k_snapshot_records(Soa s, int n, uint8_t *__restrict__ aos, const uint8_t *__restrict__ active,
                                   const uint8_t *__restrict__ is_follower, const long long *__restrict__ index_leader) {
    constexpr int ITEM = MODEL == 0 ? 228 : 316;
    constexpr int B = MODEL == 0 ? 0 : 32;          // offset of the States block (agents.py:33-60 field order)
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n || s.id[t] < 0) return;
    const int rec_i = s.id[t];
    uint8_t *rb = aos + (size_t)rec_i * ITEM;
    uint32_t *rec = reinterpret_cast<uint32_t *>(rb);
    constexpr FieldMap fmc[] = {CIRC_FIELDS(0)};
    constexpr FieldMap fmt[] = {CIRC_FIELDS(32), THREE_FIELDS};
    constexpr int NF = MODEL == 0 ? N_FIELDS_CIRC : N_FIELDS_THREE;
#pragma unroll
    for (int f = 0; f < NF; ++f) {
        const unsigned bit = MODEL == 0 ? fmc[f < N_FIELDS_CIRC ? f : 0].bit : fmt[f].bit;
        if (bit == 0u) continue;
        const int w = (MODEL == 0 ? fmc[f < N_FIELDS_CIRC ? f : 0].offset : fmt[f].offset) >> 2;
        const int plane = MODEL == 0 ? fmc[f < N_FIELDS_CIRC ? f : 0].plane : fmt[f].plane;
        const double v = s(plane, t);
        rec[w] = (uint32_t)__double2loint(v);
        rec[w + 1] = (uint32_t)__double2hiint(v);
    }
    const unsigned long long tv = (unsigned long long)s.target[t];      // target: i8 at byte B + 2
    for (int b = 0; b < 8; ++b) rb[B + 2 + b] = (uint8_t)(tv >> (8 * b));
    if (active) rb[B + 0] = active[rec_i];
    if (is_follower) {
        rb[B + 11] = is_follower[rec_i];
        const unsigned long long lv = (unsigned long long)index_leader[rec_i];
        for (int b = 0; b < 8; ++b) rb[B + 12 + b] = (uint8_t)(lv >> (8 * b));
    }
}
