/*
 * hash_oracle.c — CPU restatement of the reference's pure-PyTorch hash encoders.
 *
 * TEST INFRASTRUCTURE ONLY (tests/, smoke(), bench.py cpu_baseline / --impl reference).
 *
 * Follows /root/reference/hash_encoding.py:11-46,78-111 (HashEmbedder) and
 * /root/reference/hash_encoding_wtmk_bit.py:99-116 (message-bit HashEmbedder) operation by
 * operation in IEEE fp32, each torch elementwise op being one rounded C operation
 * (compile with -ffp-contract=off):
 *   grid_size = 1 / resolution                         hash_encoding.py:37   (box = (0,1))
 *   idx       = (int) floor(clamp(x,0,1) / grid_size)  hash_encoding.py:33-39
 *   vmin      = idx * grid_size ; vmax = vmin + grid_size          :40-41
 *   slot      = (ix*1 ^ iy*2654435761 ^ iz*805459861) & (T-1)      :11-22 (int64 there; the low
 *               log2_T bits equal the uint32 wrap-around result)
 *   corner k  = i*4 + j*2 + k (x is the MSB)                       :8
 *   w         = (x - vmin) / (vmax - vmin)   (x UNclamped)         :87
 *   trilerp along x, then y, then z with separate mul/mul/add      :89-104
 *
 * Pinned against the reference modules themselves: tests/golden/hash_*.npz were produced by
 * importing /root/reference/hash_encoding*.py on CPU (tests/golden/make_golden_hash.py).
 */
#include <math.h>
#include <stdint.h>
#include <stddef.h>

static inline float clamp01(float v) { return fminf(fmaxf(v, 0.0f), 1.0f); }

typedef struct {
    uint32_t slot[8];
    float wx, wy, wz;
} voxel_t;

static inline void locate(const float* x, float resolution, uint32_t mask, voxel_t* v) {
    const float grid_size = 1.0f / resolution;
    int32_t idx[3];
    float w[3];
    for (int a = 0; a < 3; ++a) {
        const float xc = clamp01(x[a]);
        idx[a] = (int32_t)floorf(xc / grid_size);
        const float vmin = (float)idx[a] * grid_size;
        const float vmax = vmin + grid_size;
        w[a] = (x[a] - vmin) / (vmax - vmin);
    }
    v->wx = w[0]; v->wy = w[1]; v->wz = w[2];
    for (int k = 0; k < 8; ++k) {
        const uint32_t ix = (uint32_t)(idx[0] + ((k >> 2) & 1));
        const uint32_t iy = (uint32_t)(idx[1] + ((k >> 1) & 1));
        const uint32_t iz = (uint32_t)(idx[2] + (k & 1));
        v->slot[k] = (ix ^ (iy * 2654435761u) ^ (iz * 805459861u)) & mask;
    }
}

/* hash_encoding.py:78-104 for one feature channel */
static inline float trilerp(const float e[8], float wx, float wy, float wz) {
    const float c00 = e[0] * (1 - wx) + e[4] * wx;
    const float c01 = e[1] * (1 - wx) + e[5] * wx;
    const float c10 = e[2] * (1 - wx) + e[6] * wx;
    const float c11 = e[3] * (1 - wx) + e[7] * wx;
    const float c0 = c00 * (1 - wy) + c10 * wy;
    const float c1 = c01 * (1 - wy) + c11 * wy;
    return c0 * (1 - wz) + c1 * wz;
}

/* HashEmbedder.forward, hash_encoding.py:96-111.  out [B, 2*n_levels]; slots (optional)
 * [B, n_levels, 8] int32. */
void oracle_hash_encode_forward(const float* x, uint32_t B, const float* const* tables,
                                const float* resolutions, uint32_t n_levels, uint32_t log2_T,
                                float* out, int32_t* slots) {
    const uint32_t mask = (1u << log2_T) - 1u;
#pragma omp parallel for schedule(static)
    for (int64_t b = 0; b < (int64_t)B; ++b) {
        for (uint32_t l = 0; l < n_levels; ++l) {
            voxel_t v;
            locate(x + b * 3, resolutions[l], mask, &v);
            float e0[8], e1[8];
            for (int k = 0; k < 8; ++k) {
                e0[k] = tables[l][(size_t)v.slot[k] * 2];
                e1[k] = tables[l][(size_t)v.slot[k] * 2 + 1];
                if (slots) slots[((size_t)b * n_levels + l) * 8 + k] = (int32_t)v.slot[k];
            }
            out[(size_t)b * 2 * n_levels + 2 * l] = trilerp(e0, v.wx, v.wy, v.wz);
            out[(size_t)b * 2 * n_levels + 2 * l + 1] = trilerp(e1, v.wx, v.wy, v.wz);
        }
    }
}

/* autograd of the above w.r.t. the tables: the chain rule through hash_encoding.py:89-104
 * gives corner k the weight (z-factor * y-factor * x-factor); nn.Embedding's backward
 * index-adds it.  Serial (deterministic) accumulation. grad_tables must be zero-filled. */
void oracle_hash_encode_backward(const float* x, const float* grad_out, uint32_t B,
                                 float* const* grad_tables, const float* resolutions,
                                 uint32_t n_levels, uint32_t log2_T) {
    const uint32_t mask = (1u << log2_T) - 1u;
    for (uint32_t b = 0; b < B; ++b) {
        for (uint32_t l = 0; l < n_levels; ++l) {
            voxel_t v;
            locate(x + (size_t)b * 3, resolutions[l], mask, &v);
            for (int f = 0; f < 2; ++f) {
                const float g = grad_out[(size_t)b * 2 * n_levels + 2 * l + f];
                for (int k = 0; k < 8; ++k) {
                    const float fz = (k & 1) ? v.wz : (1 - v.wz);
                    const float fy = ((k >> 1) & 1) ? v.wy : (1 - v.wy);
                    const float fx = ((k >> 2) & 1) ? v.wx : (1 - v.wx);
                    grad_tables[l][(size_t)v.slot[k] * 2 + f] += ((g * fz) * fy) * fx;
                }
            }
        }
    }
}

/* Message-bit HashEmbedder.forward, hash_encoding_wtmk_bit.py:99-116: for bit i gather from
 * table 2*i + int(message[i]) at `resolution` (2048 for every i), trilerp, and sum
 * the message_dim results (summed here in bit order; torch.sum's order differs at the 1e-7
 * level).  out [B, 2]. */
void oracle_msg_encode_forward(const float* x, uint32_t B, const float* const* tables,
                               uint32_t message_dim, const float* message, float resolution,
                               uint32_t log2_T, float* out) {
    const uint32_t mask = (1u << log2_T) - 1u;
#pragma omp parallel for schedule(static)
    for (int64_t b = 0; b < (int64_t)B; ++b) {
        voxel_t v;
        locate(x + b * 3, resolution, mask, &v);
        float acc0 = 0.f, acc1 = 0.f;
        for (uint32_t i = 0; i < message_dim; ++i) {
            const float* tab = tables[2 * i + (uint32_t)(int)message[i]];
            float e0[8], e1[8];
            for (int k = 0; k < 8; ++k) {
                e0[k] = tab[(size_t)v.slot[k] * 2];
                e1[k] = tab[(size_t)v.slot[k] * 2 + 1];
            }
            acc0 += trilerp(e0, v.wx, v.wy, v.wz);
            acc1 += trilerp(e1, v.wx, v.wy, v.wz);
        }
        out[(size_t)b * 2] = acc0;
        out[(size_t)b * 2 + 1] = acc1;
    }
}

/* gradient of the message encoder w.r.t. the pre-summed table S (every selected table's
 * gradient equals it).  G [T,2] must be zero-filled. */
void oracle_msg_encode_backward(const float* x, const float* grad_out, uint32_t B, float resolution,
                                uint32_t log2_T, float* G) {
    const float* res = &resolution;
    float* tabs[1] = {G};
    oracle_hash_encode_backward(x, grad_out, B, tabs, res, 1, log2_T);
}
