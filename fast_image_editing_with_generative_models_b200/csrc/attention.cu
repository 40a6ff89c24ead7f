// Flash-style attention for head_dim 64 on tcgen05 tensor cores (sm_100a).
// Replaces F.scaled_dot_product_attention (AttnProcessor2_0) for the UNet / ControlNet self- and cross-attention
// of the reference's edit path (diffusers call at reference src/pipeline.py:261-272).  No mask, scale 1/sqrt(d).
//
// One CTA = one 128-row Q tile of one (batch, head); it streams 128-row K/V tiles through a 2-stage TMA ring.
//   S = Q K^T      tcgen05.mma 128x128x64 -> TMEM columns [0,128)
//   softmax        4 warps, one thread per row (TMEM lane == row, so no cross-thread reductions); online max / sum;
//                  P is written as fp16 into shared memory in the K-major 128B-swizzled operand layout
//   O_j = P V_j    tcgen05.mma 128x64x128 (V consumed MN-major straight from its row-major TMA tile) -> TMEM [128,192)
//   the running output is kept in registers and rescaled by exp2(m_old - m_new) per tile.
// Two CTAs fit per SM (112 KiB smem, 256 TMEM columns each), so one CTA's softmax overlaps the other's MMAs.
// Warp roles (192 threads): w0-3 softmax/epilogue, w4 TMA producer, w5 MMA issuer + TMEM allocator.
#include "tc_common.cuh"

namespace fie {

constexpr int ATT_BM = 128, ATT_BN = 128, ATT_D = 64;
constexpr int ATT_TILE_BYTES = 128 * 64 * 2;              // 16 KiB (Q, K or V tile)
constexpr int ATT_SMEM = ATT_TILE_BYTES * 5 + 32768 + 128;    // Q + 2x(K,V) + P + barriers (2 CTAs / SM)

struct AttnParams {
    CUtensorMap q_map, k_map, v_map;
    __half* out;
    long long ldo;
    int nq, nkv;
    float scale_log2;
};

__global__ void __launch_bounds__(192, 2) k_attention_d64(const __grid_constant__ AttnParams p) {
    extern __shared__ __align__(1024) uint8_t smem[];     // no static smem in this kernel: base is 1024-aligned
    uint8_t* sQ = smem;
    uint8_t* sKV = smem + ATT_TILE_BYTES;                 // stage s: K at s*32K, V at s*32K + 16K
    uint8_t* sP = smem + ATT_TILE_BYTES * 5;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + ATT_TILE_BYTES * 5 + 32768);
    uint64_t& q_full = bars[0]; uint64_t* kv_full = bars + 1; uint64_t* kv_empty = bars + 3;
    uint64_t& s_full = bars[5]; uint64_t& p_full = bars[6]; uint64_t& o_full = bars[7];
    uint32_t& tmem_slot = *reinterpret_cast<uint32_t*>(bars + 8);
    if ((smem_u32(smem) & 1023u) != 0) { if (threadIdx.x == 0) printf("fie: attention smem base misaligned\n"); __trap(); }

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int qt = blockIdx.x, head = blockIdx.y, b = blockIdx.z;
    const int n_tiles = (p.nkv + ATT_BN - 1) / ATT_BN;

    if (threadIdx.x == 0) {
        mbar_init(&q_full, 1);
        for (int s = 0; s < 2; ++s) { mbar_init(&kv_full[s], 1); mbar_init(&kv_empty[s], 1); }
        mbar_init(&s_full, 1); mbar_init(&p_full, 128); mbar_init(&o_full, 1);
        mbar_fence_init();
    }
    if (warp == 4 && lane == 0) { tma_prefetch_desc(&p.q_map); tma_prefetch_desc(&p.k_map); tma_prefetch_desc(&p.v_map); }
    if (warp == 5) tmem_alloc(&tmem_slot, 256);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_slot;
    const uint32_t tmem_S = tmem, tmem_O = tmem + 128;

    if (warp == 4) {
        if (lane == 0) {
            mbar_arrive_expect_tx(&q_full, ATT_TILE_BYTES);
            tma_load_3d(&p.q_map, &q_full, sQ, head * ATT_D, qt * ATT_BM, b);
            for (int j = 0; j < n_tiles; ++j) {
                const int s = j & 1;
                mbar_wait(&kv_empty[s], (uint32_t)(((j >> 1) & 1) ^ 1));
                mbar_arrive_expect_tx(&kv_full[s], 2 * ATT_TILE_BYTES);
                tma_load_3d(&p.k_map, &kv_full[s], sKV + s * 2 * ATT_TILE_BYTES, head * ATT_D, j * ATT_BN, b);
                tma_load_3d(&p.v_map, &kv_full[s], sKV + s * 2 * ATT_TILE_BYTES + ATT_TILE_BYTES, head * ATT_D, j * ATT_BN, b);
            }
        }
    } else if (warp == 5) {
        const uint32_t idesc_qk = umma_idesc_f16(ATT_BM, ATT_BN, 0, 0);
        const uint32_t idesc_pv = umma_idesc_f16(ATT_BM, ATT_D, 0, 1);   // B (= V) is MN-major
        const uint32_t aQ = smem_u32(sQ), aP = smem_u32(sP);
        auto issue_qk = [&](int j) {
            const int s = j & 1;
            mbar_wait(&kv_full[s], (uint32_t)((j >> 1) & 1));
            tc_fence_after();
            if (lane == 0) {
                const uint64_t ad = umma_desc_sw128(aQ), bd = umma_desc_sw128(smem_u32(sKV + s * 2 * ATT_TILE_BYTES));
#pragma unroll
                for (int k = 0; k < ATT_D / 16; ++k) umma_f16(tmem_S, ad + 2 * k, bd + 2 * k, idesc_qk, k ? 1u : 0u);
                umma_commit(&s_full);
            }
            __syncwarp();
        };
        mbar_wait(&q_full, 0);
        issue_qk(0);
        for (int j = 0; j < n_tiles; ++j) {
            const int s = j & 1;
            mbar_wait(&p_full, (uint32_t)(j & 1));
            tc_fence_after();
            if (lane == 0) {
                const uint32_t aV = smem_u32(sKV + s * 2 * ATT_TILE_BYTES + ATT_TILE_BYTES);
#pragma unroll
                for (int k = 0; k < ATT_BN / 16; ++k) {
                    const uint64_t ad = umma_desc_sw128(aP + (k >> 2) * 16384 + (k & 3) * 32);
                    const uint64_t bd = umma_desc_sw128(aV + k * 2048, 1024, 1024);
                    umma_f16(tmem_O, ad, bd, idesc_pv, k ? 1u : 0u);
                }
                umma_commit(&kv_empty[s]);
                umma_commit(&o_full);
            }
            __syncwarp();
            if (j + 1 < n_tiles) issue_qk(j + 1);
        }
    } else {
        // ===================== softmax / epilogue: thread = row =====================
        const int row = warp * 32 + lane;
        const uint32_t lane_addr = (uint32_t)(warp * 32) << 16;
        float o_acc[ATT_D];
#pragma unroll
        for (int i = 0; i < ATT_D; ++i) o_acc[i] = 0.f;
        float m_run = -INFINITY, l_run = 0.f;
        const float sl2 = p.scale_log2;
        uint8_t* prow = sP + row * 128;
        const int sw = row & 7;
        for (int j = 0; j < n_tiles; ++j) {
            mbar_wait(&s_full, (uint32_t)(j & 1));
            tc_fence_after();
            const int kv_left = p.nkv - j * ATT_BN;     // valid columns in this tile
            float mx = m_run;
#pragma unroll 1
            for (int c = 0; c < 4; ++c) {
                uint32_t r[32];
                tmem_ld_32x32(tmem_S + lane_addr + c * 32, r);
                tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < 32; ++i) if (c * 32 + i < kv_left) mx = fmaxf(mx, __uint_as_float(r[i]));
            }
            const float alpha = exp2f((m_run - mx) * sl2);
            m_run = mx;
            if (j > 0) {
                mbar_wait(&o_full, (uint32_t)((j - 1) & 1));
                tc_fence_after();
#pragma unroll
                for (int c = 0; c < 2; ++c) {
                    uint32_t r[32];
                    tmem_ld_32x32(tmem_O + lane_addr + c * 32, r);
                    tmem_ld_wait();
#pragma unroll
                    for (int i = 0; i < 32; ++i) o_acc[c * 32 + i] = (o_acc[c * 32 + i] + __uint_as_float(r[i])) * alpha;
                }
            }
            const float mneg = -mx * sl2;
            float psum = 0.f;
#pragma unroll 1
            for (int c = 0; c < 4; ++c) {
                uint32_t r[32];
                tmem_ld_32x32(tmem_S + lane_addr + c * 32, r);
                tmem_ld_wait();
                uint32_t pk[16];
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    float p0 = (c * 32 + 2 * i < kv_left) ? exp2f(fmaf(__uint_as_float(r[2 * i]), sl2, mneg)) : 0.f;
                    float p1 = (c * 32 + 2 * i + 1 < kv_left) ? exp2f(fmaf(__uint_as_float(r[2 * i + 1]), sl2, mneg)) : 0.f;
                    psum += p0 + p1;
                    __half2 h = __floats2half2_rn(p0, p1);
                    pk[i] = *reinterpret_cast<uint32_t*>(&h);
                }
                uint8_t* base = prow + (c >> 1) * 16384;
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int unit = ((c & 1) * 4 + u) ^ sw;
                    *reinterpret_cast<uint4*>(base + unit * 16) = make_uint4(pk[4 * u], pk[4 * u + 1], pk[4 * u + 2], pk[4 * u + 3]);
                }
            }
            l_run = l_run * alpha + psum;
            fence_proxy_async_smem();
            tc_fence_before();
            mbar_arrive(&p_full);
        }
        mbar_wait(&o_full, (uint32_t)((n_tiles - 1) & 1));
        tc_fence_after();
        const float inv_l = 1.0f / l_run;
        const int qrow = qt * ATT_BM + row;
        __half* orow = p.out + ((long long)b * p.nq + qrow) * p.ldo + head * ATT_D;
#pragma unroll
        for (int c = 0; c < 2; ++c) {
            uint32_t r[32];
            tmem_ld_32x32(tmem_O + lane_addr + c * 32, r);
            tmem_ld_wait();
            if (qrow < p.nq) {
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    uint4 v; __half2* hh = reinterpret_cast<__half2*>(&v);
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const int e = u * 8 + 2 * i;
                        hh[i] = __floats2half2_rn((o_acc[c * 32 + e] + __uint_as_float(r[e])) * inv_l, (o_acc[c * 32 + e + 1] + __uint_as_float(r[e + 1])) * inv_l);
                    }
                    *reinterpret_cast<uint4*>(orow + c * 32 + u * 8) = v;
                }
            }
        }
        tc_fence_before();
    }
    __syncthreads();
    if (warp == 5) { tc_fence_after(); tmem_dealloc(tmem, 256); }
}

}  // namespace fie
using namespace fie;

extern "C" int fie_attention_d64_f16(const void* q, long long ldq, const void* k, long long ldk, const void* v, long long ldv,
                                     void* out, long long ldo, int b, int heads, int nq, int nkv, float scale, void* stream) {
    FIE_REQUIRE(q && k && v && out, "fie_attention_d64_f16: null pointer");
    FIE_REQUIRE(b > 0 && heads > 0 && nq > 0 && nkv > 0 && b <= 65535 && heads <= 65535, "fie_attention_d64_f16: bad shape");
    FIE_REQUIRE((ldq % 8) == 0 && (ldk % 8) == 0 && (ldv % 8) == 0 && (ldo % 8) == 0, "fie_attention_d64_f16: leading dims must be multiples of 8");
    FIE_REQUIRE(ldq >= heads * 64 && ldk >= heads * 64 && ldv >= heads * 64 && ldo >= heads * 64, "fie_attention_d64_f16: leading dim < heads*64");
    AttnParams p;
    memset(&p, 0, sizeof(p));
    const uint32_t box[3] = {64, 128, 1};
    int rc;
    {
        const uint64_t dims[3] = {(uint64_t)heads * 64, (uint64_t)nq, (uint64_t)b};
        const uint64_t str[2] = {(uint64_t)ldq * 2, (uint64_t)nq * ldq * 2};
        if ((rc = make_tmap_f16(&p.q_map, q, 3, dims, str, box))) return rc;
    }
    {
        const uint64_t dims[3] = {(uint64_t)heads * 64, (uint64_t)nkv, (uint64_t)b};
        const uint64_t strk[2] = {(uint64_t)ldk * 2, (uint64_t)nkv * ldk * 2};
        const uint64_t strv[2] = {(uint64_t)ldv * 2, (uint64_t)nkv * ldv * 2};
        if ((rc = make_tmap_f16(&p.k_map, k, 3, dims, strk, box))) return rc;
        if ((rc = make_tmap_f16(&p.v_map, v, 3, dims, strv, box))) return rc;
    }
    p.out = (__half*)out; p.ldo = ldo; p.nq = nq; p.nkv = nkv;
    p.scale_log2 = scale * 1.4426950408889634f;
    static bool attr = false;
    if (!attr) {
        cudaError_t e = cudaFuncSetAttribute(k_attention_d64, cudaFuncAttributeMaxDynamicSharedMemorySize, ATT_SMEM);
        if (e != cudaSuccess) { set_error("cudaFuncSetAttribute(k_attention_d64): %s", cudaGetErrorString(e)); return FIE_ERR_CUDA; }
        attr = true;
    }
    dim3 grid((nq + ATT_BM - 1) / ATT_BM, heads, b);
    k_attention_d64<<<grid, 192, ATT_SMEM, (cudaStream_t)stream>>>(p);
    return check_launch("fie_attention_d64_f16");
}
