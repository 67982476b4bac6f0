// TEST INFRASTRUCTURE -- NOT part of the product path.
//
// Host lock-step emulator for the warp body of the 16-bit extension kernel
// (hite_b200/csrc/hb_extend16.cuh): 32 host threads play the 32 lanes of one warp, every shuffle /
// warp-sync is a barrier exchange.  It lets the CPU test-suite (which has no GPU) run the very same
// source the device compiles -- band geometry, packed s16x2 recurrences, traceback nibble packing,
// stream layout, best-cell reduction -- against the oracle (oracle/hbo_align.c::hbo_align_pair).
// Built on demand by tests/test_extend16_emu.py with `nvcc -x cu` (host code only is executed).
#include <atomic>
#include <cstring>
#include <thread>
#include <vector>

#include "../../hite_b200/csrc/hb_extend16.cuh"

namespace {

struct SpinBarrier {
    std::atomic<int> count{0};
    std::atomic<int> sense{0};
    int n = 32;
    void wait() {
        const int s = sense.load(std::memory_order_acquire);
        if (count.fetch_add(1, std::memory_order_acq_rel) == n - 1) {
            count.store(0, std::memory_order_relaxed);
            sense.store(s ^ 1, std::memory_order_release);
        } else {
            int spins = 0;
            while (sense.load(std::memory_order_acquire) == s)
                if (++spins > 64) { std::this_thread::yield(); spins = 0; }
        }
    }
};

struct WarpShared {
    SpinBarrier bar;
    unsigned long long slot[32];
};

struct EmuWarp {
    WarpShared* sh;
    int lane_;
    int lane() const { return lane_; }
    void sync() const { sh->bar.wait(); }
    unsigned long long exchange(unsigned long long v, int src) const {
        sh->slot[lane_] = v;
        sh->bar.wait();
        const unsigned long long r = sh->slot[src];
        sh->bar.wait();
        return r;
    }
    uint32_t shfl_up(uint32_t v, int d, int width) const {
        const int seg = lane_ / width * width;
        const int src = (lane_ - d >= seg) ? lane_ - d : lane_;
        return (uint32_t)exchange(v, src);
    }
    uint32_t shfl_down(uint32_t v, int d, int width) const {
        const int seg = lane_ / width * width;
        const int src = (lane_ + d < seg + width) ? lane_ + d : lane_;
        return (uint32_t)exchange(v, src);
    }
    unsigned long long shfl_xor64(unsigned long long v, int off) const { return exchange(v, lane_ ^ off); }
    int max_all(int v) const {
        sh->slot[lane_] = (unsigned long long)(long long)v;
        sh->bar.wait();
        int m = v;
        for (int l = 0; l < 32; ++l) { const int x = (int)(long long)sh->slot[l]; if (x > m) m = x; }
        sh->bar.wait();
        return m;
    }
};

void pack_codes(const uint8_t* codes, uint64_t n, std::vector<uint32_t>& seq2, std::vector<uint32_t>& nmask) {
    seq2.assign((n + 15) / 16 + 2, 0u);
    nmask.assign((n + 31) / 32 + 2, 0u);
    for (uint64_t p = 0; p < n; ++p) {
        if (codes[p] > 3) nmask[p >> 5] |= 1u << (p & 31);
        else seq2[p >> 4] |= (uint32_t)codes[p] << (2 * (p & 15));
    }
}

template <int HB>
void run_class(const std::vector<uint32_t>& order, const HbLocus* pairs, const uint32_t* gseq2, const uint32_t* gnmask,
               uint64_t tot, const uint32_t* cseq2, const uint32_t* cnmask, const uint64_t* cgoff, const uint32_t* clen,
               const uint32_t* tb_off, uint32_t* tbuf, HbAln* alns) {
    constexpr int PPW = 32 / HB, WIN = HB + 32;
    const uint32_t n_order = (uint32_t)order.size();
    const uint32_t n_groups = (n_order + 1) / 2;
    for (uint32_t group0 = 0; group0 < n_groups; group0 += PPW) {
        WarpShared sh;
        std::vector<uint32_t> s_rq(PPW * WIN, 0u), s_rt(PPW * WIN, 0u);
        std::vector<std::thread> th;
        for (int l = 0; l < 32; ++l)
            th.emplace_back([&, l]() {
                EmuWarp w{&sh, l};
                hb16::extend16_warp<HB>(w, group0, order.data(), n_order, pairs, gseq2, gnmask, tot, cseq2, cnmask, cgoff,
                                        clen, tb_off, 0u, tbuf, alns, s_rq.data(), s_rt.data());
            });
        for (auto& t : th) t.join();
    }
}

}  // namespace

// genome_codes: tot entries 0..4 (the oracle's padded layout); cand_codes: concatenated candidates.
// pairs: cand, rel, j0, band class (0 wide = 64 diagonals, 1 narrow = 32, 2 tight = 16) per pair.  Outputs per pair: score, i_s, i_e, j_s, j_e and qrec rows
// (qrec_off[p] .. + L words, zero outside the alignment).
extern "C" int emu_extend16(const uint8_t* genome_codes, uint64_t tot, const uint8_t* cand_codes, const uint64_t* cand_off,
                            uint32_t n_cand, const uint32_t* p_cand, const uint32_t* p_rel, const int64_t* p_j0,
                            const uint32_t* p_narrow, uint32_t n_pairs, int32_t* out_score, int32_t* out_is,
                            int32_t* out_ie, int64_t* out_js, int64_t* out_je, uint32_t* out_qrec,
                            const uint64_t* out_qrec_off) {
    std::vector<uint32_t> gseq2, gnmask, cseq2, cnmask;
    pack_codes(genome_codes, tot, gseq2, gnmask);
    // candidates in one padded space like the device (64 N before / between)
    std::vector<uint64_t> cgoff(n_cand + 1);
    std::vector<uint32_t> clen(n_cand);
    uint64_t p = 64;
    for (uint32_t c = 0; c < n_cand; ++c) { cgoff[c] = p; clen[c] = (uint32_t)(cand_off[c + 1] - cand_off[c]); p += clen[c] + 64; }
    const uint64_t ctot = (p + 63) & ~63ull;
    std::vector<uint8_t> cc(ctot, 4);
    for (uint32_t c = 0; c < n_cand; ++c) memcpy(cc.data() + cgoff[c], cand_codes + cand_off[c], clen[c]);
    pack_codes(cc.data(), ctot, cseq2, cnmask);

    std::vector<HbLocus> pairs(n_pairs);
    std::vector<uint32_t> tb_off(n_pairs + 1), qr_off(n_pairs + 1), order_t, order_n, order_w;
    uint32_t tbu = 0, qru = 0;
    for (uint32_t i = 0; i < n_pairs; ++i) {
        HbLocus& pr = pairs[i];
        pr.cand = p_cand[i]; pr.rel = p_rel[i]; pr.nseeds = 3; pr.rank = i; pr.j0 = p_j0[i];
        pr.dmin = 0; pr.dmax = p_narrow[i] == 2 ? 0 : p_narrow[i] == 1 ? hb16::NARROW_SPREAD : 100;   // band class: 0 wide, 1 narrow, 2 tight
        const uint32_t L = clen[pr.cand];
        if ((int)L > hb16::HB16_MAX_L) return -1;
        const uint32_t hb = (uint32_t)hb16::half_band(pr.dmax - pr.dmin, (int)L);
        tb_off[i] = tbu; qr_off[i] = qru;
        tbu += (((2 * L + 2 * hb - 2 + 63) >> 6) * hb + 15) / 16;
        qru += (L + 31) >> 5;
        (hb == 8 ? order_t : hb == 16 ? order_n : order_w).push_back(i);
    }
    std::vector<uint32_t> tbuf(((size_t)tbu + 1) << 7, 0xDEADBEEFu);       // poison: unwritten sectors must never be read
    std::vector<uint32_t> qrec(((size_t)qru + 1) << 5, 0u);
    std::vector<HbAln> alns(n_pairs);
    memset(alns.data(), 0xFF, sizeof(HbAln) * n_pairs);
    run_class<32>(order_w, pairs.data(), gseq2.data(), gnmask.data(), tot, cseq2.data(), cnmask.data(), cgoff.data(),
                  clen.data(), tb_off.data(), tbuf.data(), alns.data());
    run_class<16>(order_n, pairs.data(), gseq2.data(), gnmask.data(), tot, cseq2.data(), cnmask.data(), cgoff.data(),
                  clen.data(), tb_off.data(), tbuf.data(), alns.data());
    run_class<8>(order_t, pairs.data(), gseq2.data(), gnmask.data(), tot, cseq2.data(), cnmask.data(), cgoff.data(),
                 clen.data(), tb_off.data(), tbuf.data(), alns.data());
    for (uint32_t i = 0; i < n_pairs; ++i)
        hb16::traceback_pair(i, pairs.data(), clen.data(), tb_off.data(), 0u, tbuf.data(), qr_off.data(), qrec.data(), alns.data());
    for (uint32_t i = 0; i < n_pairs; ++i) {
        out_score[i] = alns[i].score; out_is[i] = alns[i].i_s; out_ie[i] = alns[i].i_e;
        out_js[i] = alns[i].j_s; out_je[i] = alns[i].j_e;
        const uint32_t L = clen[pairs[i].cand];
        memcpy(out_qrec + out_qrec_off[i], qrec.data() + ((size_t)qr_off[i] << 5), sizeof(uint32_t) * L);
    }
    return 0;
}
