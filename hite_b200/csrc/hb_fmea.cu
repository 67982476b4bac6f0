// hite_b200 -- FMEA interval chaining of the coarse-boundary stage (host C++, inside the C-ABI library).
//
// Replaces get_longest_repeats_v4 (reference module/Util.py:4122-4400) with process_all_seqs
// (module/Util.py:4529-4563): per query segment, HSPs are grouped by subject and strand, clustered
// along the subject, greedily chained along the query within `fixed_extend_base_threshold`, the
// resulting query intervals are de-duplicated on 10-bp-rounded positions and on >= 95 % overlap.
// Why host code: the chaining is a data-dependent greedy walk over tiny sorted lists whose
// de-duplication is order-dependent (a Python dict filled in visiting order); there is no data
// parallelism inside a query segment, so query segments are the parallel grain (std::thread pool here,
// ProcessPoolExecutor in the reference, module/Util.py:4773).  The HSPs themselves come from the GPU
// (hb_coarse.cu).  Results are bit-identical to the reference on the same HSP list (tests/golden/fmea.json.gz).
#include <algorithm>
#include <atomic>
#include <chrono>
#include <climits>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <functional>
#include <condition_variable>
#include <deque>
#include <memory>
#include <mutex>
#include <thread>
#include <unordered_set>
#include <vector>

#include "hb_fmea.cuh"

namespace {

struct Frag { int32_t qs, qe, ss, se; };

struct Chain { int32_t qs, qe, ss, se, subj; };

// greedy extension inside one subject cluster (module/Util.py:4229-4319)
struct Pt { int32_t chr, da, db, used; };     // one stored point per interval of the de-duplication (see dedup_query)
// Per-worker scratch, reused across work items: repeated large malloc/free from many threads serialises on
// the kernel's address-space lock (mmap/munmap), which made the pool anti-scale.
struct Scratch {
    std::vector<Frag> sorted, cl;
    std::vector<uint64_t> keys, tmp;
    std::vector<char> used;
    std::vector<Chain> chains;
    std::vector<Pt> table;
};

// `src[0 .. m)`: one cluster, a contiguous run of the strand in subject order; its (qs, qe)-sorted copy is built in `c`
void chain_cluster(const Frag* src, size_t m, int skip_gap, int subj, std::vector<Chain>& out, std::vector<char>& used,
                   std::vector<uint64_t>& keys, std::vector<Frag>& c) {
    if (m == 1) { out.push_back(Chain{src[0].qs, src[0].qe, src[0].ss, src[0].se, subj}); return; }   // the common case
    {   // stable sort on (qs, qe) (:4231): one packed key per fragment when the coordinates fit, as in sort_strand
        bool fits = m < (1u << 22);
        keys.resize(m);
        for (size_t i = 0; fits && i < m; ++i) {
            const uint32_t qs = (uint32_t)src[i].qs, qe = (uint32_t)src[i].qe;
            if ((qs | qe) >> 21) fits = false;
            keys[i] = (uint64_t)qs << 43 | (uint64_t)qe << 22 | (uint64_t)i;
        }
        c.resize(m);
        if (fits) {
            std::sort(keys.begin(), keys.end());
            for (size_t i = 0; i < m; ++i) c[i] = src[keys[i] & 0x3FFFFFu];
        } else {
            for (size_t i = 0; i < m; ++i) c[i] = src[i];
            std::stable_sort(c.begin(), c.end(), [](const Frag& a, const Frag& b) { return a.qs != b.qs ? a.qs < b.qs : a.qe < b.qe; });
        }
    }
    const size_t n = c.size();
    used.assign(n, 0);
    // visited_frag is keyed by the HSP tuple: identical lines fall together.  After the sort they can only sit in the run of
    // equal (qs, qe) around i, so that run is all there is to look at.
    auto mark = [&](size_t i) {
        const int32_t qs = c[i].qs, qe = c[i].qe, ss = c[i].ss, se = c[i].se;
        used[i] = 1;
        for (size_t t = i; t-- > 0 && c[t].qs == qs && c[t].qe == qe;) if (c[t].ss == ss && c[t].se == se) used[t] = 1;
        for (size_t t = i + 1; t < n && c[t].qs == qs && c[t].qe == qe; ++t) if (c[t].ss == ss && c[t].se == se) used[t] = 1;
    };
    for (size_t i = 0; i < n; ++i) {
        if (used[i]) continue;
        Chain ch{c[i].qs, c[i].qe, c[i].ss, c[i].se, subj};
        mark(i);
        for (size_t j = i + 1; j < n; ++j) {
            if (used[j]) continue;
            const Frag& f = c[j];
            if (f.qe <= ch.qe) continue;
            const bool fwd = ch.ss < ch.se && f.ss < f.se, rev = ch.ss > ch.se && f.ss > f.se;
            if (fwd && f.se > ch.se) {
                if (f.qs - ch.qe < skip_gap && f.ss - ch.se < skip_gap) {
                    ch.qe = f.qe; ch.ss = std::min(ch.ss, f.ss); ch.se = f.se;
                    mark(j);
                } else if (f.qs - ch.qe >= skip_gap) break;
            } else if (rev && f.se < ch.se) {
                if (f.qs - ch.qe < skip_gap && ch.se - f.ss < skip_gap) {
                    ch.qe = f.qe; ch.ss = std::max(ch.ss, f.ss); ch.se = f.se;
                    mark(j);
                } else if (f.qs - ch.qe >= skip_gap) break;
            }
        }
        out.push_back(ch);
    }
}

// Stable LSD radix sort of packed 64-bit keys on the bit range [lo, hi): 11-bit digits, counting passes; a digit on which
// all keys agree costs one histogram pass and no scatter.
void radix_sort_keys(std::vector<uint64_t>& keys, std::vector<uint64_t>& tmp, int lo, int hi) {
    const size_t n = keys.size();
    tmp.resize(n);
    uint64_t* a = keys.data();
    uint64_t* b = tmp.data();
    uint32_t hist[2048];
    for (int sh = lo; sh < hi; sh += 11) {
        const int bits = std::min(11, hi - sh);
        const uint32_t mask = (1u << bits) - 1u;
        for (uint32_t d = 0; d <= mask; ++d) hist[d] = 0;
        for (size_t i = 0; i < n; ++i) hist[(a[i] >> sh) & mask]++;
        if (hist[(a[0] >> sh) & mask] == n) continue;
        uint32_t run = 0;
        for (uint32_t d = 0; d <= mask; ++d) { const uint32_t c = hist[d]; hist[d] = run; run += c; }
        for (size_t i = 0; i < n; ++i) b[hist[(a[i] >> sh) & mask]++] = a[i];
        std::swap(a, b);
    }
    if (a != keys.data()) keys.swap(tmp);
}

// Stable sort of one strand's fragments along the subject (module/Util.py:4174-4176): forward ascending on (ss, se), reverse
// descending.  Coordinates are 1-based positions inside a segment; when they (and the count) fit 21 / 21 / 22 bits the order
// is the order of one packed 64-bit key per fragment (position in file order in the low bits keeps it stable), otherwise a
// comparison sort on the records.  `out` only ever grows (no value-initialisation per call); the strand is out[0 .. n).
void sort_strand(const Frag* f, size_t n, bool reverse, std::vector<Frag>& out, std::vector<uint64_t>& keys, std::vector<uint64_t>& tmp) {
    if (out.size() < n) out.resize(n + n / 4);
    if (n == 0) return;
    if (n == 1) { out[0] = f[0]; return; }
    bool fits = n < (1u << 22);
    if (fits) {
        keys.resize(n);
        for (size_t i = 0; i < n; ++i) {
            const uint32_t ss = (uint32_t)f[i].ss, se = (uint32_t)f[i].se;
            if ((ss | se) >> 21) { fits = false; break; }
            const uint64_t a = reverse ? 0x1FFFFFu - ss : ss, b = reverse ? 0x1FFFFFu - se : se;
            keys[i] = a << 43 | b << 22 | (uint64_t)i;
        }
    }
    if (fits) {
        if (n < 192) std::sort(keys.begin(), keys.end());
        else {
            // two counting passes on the subject start alone (stable), then the few runs of equal starts ordered on (end, file
            // position) by the rest of the key
            radix_sort_keys(keys, tmp, 43, 64);
            for (size_t a = 0; a < n;) {
                size_t b = a + 1;
                while (b < n && (keys[b] >> 43) == (keys[a] >> 43)) ++b;
                if (b - a > 1) std::sort(keys.begin() + (ptrdiff_t)a, keys.begin() + (ptrdiff_t)b);
                a = b;
            }
        }
        for (size_t i = 0; i < n; ++i) out[i] = f[keys[i] & 0x3FFFFFu];
        return;
    }
    for (size_t i = 0; i < n; ++i) out[i] = f[i];
    if (!reverse)
        std::stable_sort(out.begin(), out.begin() + (ptrdiff_t)n, [](const Frag& a, const Frag& b) { return a.ss != b.ss ? a.ss < b.ss : a.se < b.se; });
    else
        std::stable_sort(out.begin(), out.begin() + (ptrdiff_t)n, [](const Frag& a, const Frag& b) { return a.ss != b.ss ? a.ss > b.ss : a.se > b.se; });
}

// chains of one (query, subject) pair: forward clusters, then reverse clusters (module/Util.py:4162-4319).
// `fw` / `rv`: the pair's HSPs of either strand in file order (:4168-4172).
void chain_subject(const Frag* fw, size_t n_fw, const Frag* rv, size_t n_rv, int skip_gap, int subj, std::vector<Chain>& chains,
                   Scratch& S) {
    for (int pass = 0; pass < 2; ++pass) {                                            // :4178-4227
        const size_t np = pass ? n_rv : n_fw;
        sort_strand(pass ? rv : fw, np, pass != 0, S.sorted, S.keys, S.tmp);
        const Frag* p = S.sorted.data();
        size_t k0 = 0;                                   // the open cluster is the run p[k0 .. k)
        for (size_t k = 1; k < np; ++k) {
            bool closed = false;
            for (size_t e = k; e-- > k0;) {
                const int gap = pass ? (p[e].se - p[k].ss) : (p[k].ss - p[e].se);
                if (gap < skip_gap && p[k].qe > p[e].qe) { closed = true; break; }
            }
            if (!closed) { chain_cluster(p + k0, k - k0, skip_gap, subj, chains, S.used, S.keys, S.cl); k0 = k; }
        }
        if (np) chain_cluster(p + k0, np - k0, skip_gap, subj, chains, S.used, S.keys, S.cl);
    }
}

// per query segment: subjects in order of first appearance (Python dict order, module/Util.py:4140-4147); the HSPs of the
// query in ONE array, bucketed by (subject slot, strand) with the file order kept inside a bucket (a counting sort: two
// passes over the query's HSPs, no per-subject vectors)
struct QueryWork {
    std::vector<int32_t> order;                  // subject ids in first-appearance order
    std::vector<uint32_t> start;                 // bucket 2 * slot (forward) / 2 * slot + 1 (reverse) = frags[start[b] .. start[b + 1])
    std::vector<Frag> frags;
    std::vector<std::vector<Chain>> chains;      // filled by the (query, subject) work items
    void swap_storage(QueryWork& o) { order.swap(o.order); start.swap(o.start); frags.swap(o.frags); chains.swap(o.chains); }
};

void group_query(const hb_hsp* h, size_t n, int32_t q_index, QueryWork& w, uint32_t n_subj) {
    std::vector<int32_t> slot((size_t)n_subj, -1);
    std::vector<uint32_t>& start = w.start;
    start.clear();
    auto skip = [&](const hb_hsp& x) { return x.subj == q_index && x.qs == x.ss && x.qe == x.se; };            // :4138
    for (size_t i = 0; i < n; ++i) {
        const hb_hsp& x = h[i];
        if (skip(x)) continue;
        if ((size_t)x.subj >= slot.size()) slot.resize((size_t)x.subj + 1, -1);
        if (slot[x.subj] < 0) { slot[x.subj] = (int32_t)w.order.size(); w.order.push_back(x.subj); start.push_back(0); start.push_back(0); }
        start[2 * (size_t)slot[x.subj] + (x.ss > x.se ? 1 : 0)]++;
    }
    uint32_t run = 0;
    for (uint32_t& c : start) { const uint32_t k = c; c = run; run += k; }
    start.push_back(run);
    w.frags.resize(run);
    std::vector<uint32_t> cur(start.begin(), start.end() - 1);
    for (size_t i = 0; i < n; ++i) {
        const hb_hsp& x = h[i];
        if (skip(x)) continue;
        w.frags[cur[2 * (size_t)slot[x.subj] + (x.ss > x.se ? 1 : 0)]++] = Frag{x.qs, x.qe, x.ss, x.se};
    }
    w.chains.resize(w.order.size());
}

// One point per interval: (chromosome, start decade, end decade).  The reference rounds every interval end down and up to a
// multiple of 10 and stores all four (start, end) combinations (module/Util.py:4330-4390); two intervals share a key iff
// their start decades differ by at most one AND their end decades differ by at most one, so one stored point and a look-up of
// the nine neighbouring decade pairs answers the same question with a table a quarter of the size.
// `file_chains`: chains of ALL the query records that share this query's BLAST result file (the reference creates
// chr_pos_candidates once per file, module/Util.py:4155); `first_of_file`: this is the file's first query, the table starts empty.
void dedup_query(const QueryWork& w, const hb_segment& q, int32_t q_index, const hb_segment* segs, int max_len,
                 std::vector<hb_interval>& out, Scratch& S, size_t file_chains, bool first_of_file) {
    std::vector<Chain>& chains = S.chains;
    chains.clear();
    for (const auto& c : w.chains) chains.insert(chains.end(), c.begin(), c.end());
    // 10-bp-rounded position de-duplication in visiting order; the set is a flat open-addressing table (2 points per chain,
    // no per-node allocation)
    size_t cap = 64;
    while (cap < file_chains * 4 + 16) cap <<= 1;
    std::vector<Pt>& table = S.table;
    if (first_of_file) table.assign(cap, Pt{0, 0, 0, 0});
    else cap = table.size();
    // hashed on (chromosome, start decade) only: the three end decades of a neighbourhood sit in one probe run
    auto slot_of = [&](int32_t chr, int32_t da) -> size_t {
        uint64_t h = ((uint64_t)(uint32_t)chr << 32 | (uint32_t)da) * 0x9E3779B97F4A7C15ull;
        h ^= h >> 31;
        return (size_t)h & (cap - 1);
    };
    auto insert = [&](int32_t chr, int32_t da, int32_t db) {
        size_t i = slot_of(chr, da);
        while (table[i].used) {
            if (table[i].chr == chr && table[i].da == da && table[i].db == db) return;
            i = (i + 1) & (cap - 1);
        }
        table[i] = Pt{chr, da, db, 1};
    };
    auto near = [&](int32_t chr, int32_t da, int32_t db) -> bool {
        for (int x = -1; x <= 1; ++x) {
            size_t i = slot_of(chr, da + x);
            while (table[i].used) {
                if (table[i].chr == chr && table[i].da == da + x && (uint32_t)(table[i].db - db + 1) <= 2u) return true;
                i = (i + 1) & (cap - 1);
            }
        }
        return false;
    };
    struct Iv { int64_t a, b; };
    std::vector<Iv> names;
    const size_t n_ch = chains.size();
    for (size_t ci = 0; ci < n_ch; ++ci) {
        if (ci + 8 < n_ch) {                          // the table is L2-sized and the probes are random: fetch ahead
            const Chain& nx = chains[ci + 8];
            const hb_segment& ns = segs[nx.subj];
            const int32_t nsd = (int32_t)((ns.start + (nx.ss - 1)) / 10), nqd = (int32_t)((q.start + (nx.qs - 1)) / 10);
            for (int x = -1; x <= 1; ++x) {
                __builtin_prefetch(&table[slot_of(ns.chr, nsd + x)]);
                __builtin_prefetch(&table[slot_of(q.chr, nqd + x)]);
            }
        }
        const Chain& ch = chains[ci];
        const hb_segment& s = segs[ch.subj];
        const int64_t sa = s.start + (ch.ss - 1), sb = s.start + ch.se;
        const int64_t qa = q.start + (ch.qs - 1), qb = q.start + ch.qe;
        const int32_t sda = (int32_t)(sa / 10), sdb = (int32_t)(sb / 10), qda = (int32_t)(qa / 10), qdb = (int32_t)(qb / 10);
        const bool hit = near(s.chr, sda, sdb) || near(q.chr, qda, qdb);
        const int64_t qlen = std::llabs((long long)ch.qe - (ch.qs - 1));
        if (!hit && qlen >= 80 && qlen < max_len) names.push_back(Iv{qa, qb});
        insert(s.chr, sda, sdb);
        insert(q.chr, qda, qdb);
    }
    // >= 95 % overlap merge, longest first (module/Util.py:4529-4549): interval i removes every later (not longer) interval j
    // with overlap / len_j >= 0.95.  Such a j starts inside [a_i - 0.05 len_j, b_i), so only the intervals whose start lies in
    // [a_i - 0.06 len_i - 1, b_i) are looked at (found in a start-sorted index) -- the predicate itself is unchanged.
    // both orders through packed 64-bit keys (value, position) when the values fit -- they do: lengths are below max_len and the
    // starts lie inside one query segment
    const size_t nn = names.size();
    std::vector<uint64_t>& keys = S.keys;
    int64_t a_min = INT64_MAX, a_max = INT64_MIN, len_max = 0;
    for (const Iv& x : names) { a_min = std::min(a_min, x.a); a_max = std::max(a_max, x.a); len_max = std::max(len_max, x.b - x.a); }
    const bool packed = nn > 1 && nn < (1u << 24) && len_max < (1ll << 39) && a_max - a_min < (1ll << 39);
    if (packed) {
        keys.resize(nn);
        for (size_t i = 0; i < nn; ++i) keys[i] = (uint64_t)(len_max - (names[i].b - names[i].a)) << 24 | (uint64_t)i;
        std::sort(keys.begin(), keys.end());                                    // longest first, ties in visiting order
        std::vector<Iv> sorted(nn);
        for (size_t i = 0; i < nn; ++i) sorted[i] = names[keys[i] & 0xFFFFFFu];
        names.swap(sorted);
    } else {
        std::stable_sort(names.begin(), names.end(), [](const Iv& x, const Iv& y) { return (x.b - x.a) > (y.b - y.a); });
    }
    std::vector<uint32_t> by_start(nn);
    if (packed) {
        for (size_t i = 0; i < nn; ++i) keys[i] = (uint64_t)(names[i].a - a_min) << 24 | (uint64_t)i;
        std::sort(keys.begin(), keys.end());
        for (size_t i = 0; i < nn; ++i) by_start[i] = (uint32_t)(keys[i] & 0xFFFFFFu);
    } else {
        for (size_t i = 0; i < nn; ++i) by_start[i] = (uint32_t)i;
        std::sort(by_start.begin(), by_start.end(), [&](uint32_t x, uint32_t y) { return names[x].a != names[y].a ? names[x].a < names[y].a : x < y; });
    }
    std::vector<char> keep(nn, 1);
    for (size_t i = 0; i < nn; ++i) {
        if (!keep[i]) continue;
        const int64_t len_i = names[i].b - names[i].a;
        const int64_t lo = names[i].a - (len_i * 6) / 100 - 1;
        size_t p = (size_t)(std::lower_bound(by_start.begin(), by_start.end(), lo, [&](uint32_t x, int64_t v) { return names[x].a < v; }) - by_start.begin());
        for (; p < nn && names[by_start[p]].a < names[i].b; ++p) {
            const size_t j = by_start[p];
            if (j <= i || !keep[j]) continue;
            const int64_t ov = std::max<int64_t>(0, std::min(names[i].b, names[j].b) - std::max(names[i].a, names[j].a));
            if ((double)ov / (double)(names[j].b - names[j].a) >= 0.95) keep[j] = 0;
        }
    }
    // a name can only repeat among the kept intervals of this query (dict keys, :4551-4560)
    std::vector<std::pair<int64_t, int64_t>> seen;
    for (size_t i = 0; i < nn; ++i) if (keep[i]) seen.emplace_back(names[i].a, names[i].b);
    std::sort(seen.begin(), seen.end());
    const bool has_dup = std::adjacent_find(seen.begin(), seen.end()) != seen.end();
    for (size_t i = 0; i < nn; ++i) {
        if (!keep[i]) continue;
        bool dup = false;
        if (has_dup) for (const hb_interval& o : out) if (o.start == names[i].a && o.end == names[i].b) { dup = true; break; }
        if (!dup) out.push_back(hb_interval{q.chr, q_index, names[i].a, names[i].b});
    }
}

}  // namespace

// ------------------------------------------------------------------------------------------------
// The chaining as a task graph on one pool of host threads (no phase barriers):
//   G(q)    group the HSPs of query segment q by (subject, strand)          -> releases the C(q, k) of its subjects
//   C(q, k) chain one (query, subject) pair (the bulk of the work, ~10^4 items for a 100-segment chunk)
//   D(f)    order-dependent de-duplication of BLAST result file f, once every C of the file's queries is done
// Blocks of whole files are submitted as the GPU delivers them (hb_coarse_boundary) and are chained while the GPU produces the
// next ones; a file's D runs while the C items of later files still keep the other threads busy.
struct HbFmeaPool {
    struct Task { uint8_t type; uint32_t a, b; };       // 0 = G(a), 1 = C(a, slot b), 2 = D(file a)
    uint32_t n_seg = 0;
    const hb_segment* segs = nullptr;
    int skip_gap = 0, max_len = 0;
    std::vector<QueryWork> work;
    std::vector<std::vector<hb_interval>> per;
    std::vector<const hb_hsp*> q_hsp;                    // HSP slice of every query segment (set by submit)
    std::vector<uint64_t> q_n;
    std::vector<std::pair<uint32_t, uint32_t>> files;    // [first, last) segment of every file
    std::vector<uint32_t> file_of;                       // file index of every segment
    std::unique_ptr<std::atomic<int64_t>[]> pending;     // per file: G / C tasks not finished yet
    std::deque<Task> queue;
    std::mutex mu;
    std::condition_variable cv;
    uint64_t inflight = 0;                               // tasks pushed and not finished (under mu)
    bool closing = false;
    std::vector<std::thread> threads;
    std::vector<Scratch> scratch;
    std::atomic<int64_t> busy_us{0};
    std::chrono::steady_clock::time_point t_first{}, t_last{};
    bool any = false;

    void push_locked(const Task& t, bool front) {
        if (front) queue.push_front(t); else queue.push_back(t);
        ++inflight;
    }
    void run_task(const Task& t, Scratch& S) {
        if (t.type == 0) {
            const uint32_t g = t.a;
            group_query(q_hsp[g], (size_t)q_n[g], (int32_t)g, work[g], n_seg);
            const uint32_t n_items = (uint32_t)work[g].order.size();
            std::atomic<int64_t>& pend = pending[file_of[g]];
            pend.fetch_add((int64_t)n_items);
            if (n_items) {
                std::lock_guard<std::mutex> lk(mu);
                for (uint32_t k = 0; k < n_items; ++k) push_locked(Task{1, g, k}, false);
            }
            if (n_items > 1) cv.notify_all(); else if (n_items) cv.notify_one();
            release(file_of[g]);
        } else if (t.type == 1) {
            QueryWork& w = work[t.a];
            const uint32_t k = t.b;
            const Frag* f = w.frags.data();
            const uint32_t a = w.start[2 * k], m = w.start[2 * k + 1], b = w.start[2 * k + 2];
            chain_subject(f + a, m - a, f + m, b - m, skip_gap, w.order[k], w.chains[k], S);
            release(file_of[t.a]);
        } else {
            dedup_file(t.a, S);
        }
    }
    void release(uint32_t f) {
        if (pending[f].fetch_sub(1) == 1) {              // the file's last G / C: its de-duplication can run (long task: front)
            { std::lock_guard<std::mutex> lk(mu); push_locked(Task{2, f, 0}, true); }
            cv.notify_one();
        }
    }
    // de-duplication per BLAST result FILE: the query records of one file share the rounded-position table and the file's
    // name dict, in record order
    void dedup_file(uint32_t f, Scratch& S) {
        const uint32_t g0 = files[f].first, g1 = files[f].second;
        size_t file_chains = 0;
        for (uint32_t g = g0; g < g1; ++g) for (const auto& c : work[g].chains) file_chains += c.size();
        bool first = true;
        for (uint32_t g = g0; g < g1; ++g) {
            if (work[g].order.empty()) continue;
            dedup_query(work[g], segs[g], (int32_t)g, segs, max_len, per[g], S, file_chains, first);
            first = false;
            // longest_repeats is one dict per file (:4152): a name an earlier record of the file produced is not repeated
            if (g > g0) {
                auto& mine = per[g];
                mine.erase(std::remove_if(mine.begin(), mine.end(), [&](const hb_interval& x) {
                    for (uint32_t p = g0; p < g; ++p)
                        for (const hb_interval& o : per[p]) if (o.chr == x.chr && o.start == x.start && o.end == x.end) return true;
                    return false;
                }), mine.end());
            }
            QueryWork().swap_storage(work[g]);           // the query's fragments and chains are not needed any more
        }
    }
    void worker(int tid) {
        Scratch& S = scratch[(size_t)tid];
        for (;;) {
            Task t;
            {
                std::unique_lock<std::mutex> lk(mu);
                cv.wait(lk, [&] { return !queue.empty() || (closing && inflight == 0); });
                if (queue.empty()) return;
                t = queue.front();
                queue.pop_front();
            }
            const auto b0 = std::chrono::steady_clock::now();
            run_task(t, S);
            busy_us += (int64_t)std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - b0).count();
            bool done = false;
            {
                std::lock_guard<std::mutex> lk(mu);
                done = (--inflight == 0);
                if (done) t_last = std::chrono::steady_clock::now();
            }
            if (done) cv.notify_all();                   // finish() and the idle workers re-check
        }
    }
};

HbFmeaPool* hb_fmea_pool_start(uint32_t n_seg, const hb_segment* segs, const uint32_t* file_of_seg, int skip_gap, int max_len,
                               int threads) {
    HbFmeaPool* P = new HbFmeaPool();
    P->n_seg = n_seg; P->segs = segs; P->skip_gap = skip_gap; P->max_len = max_len;
    P->work.resize(n_seg); P->per.resize(n_seg); P->q_hsp.assign(n_seg, nullptr); P->q_n.assign(n_seg, 0);
    P->file_of.resize(n_seg);
    for (uint32_t g = 0; g < n_seg;) {                   // file_of_seg: non-decreasing file id per segment; NULL = one record per file
        uint32_t e = g + 1;
        while (file_of_seg && e < n_seg && file_of_seg[e] == file_of_seg[g]) ++e;
        for (uint32_t x = g; x < e; ++x) P->file_of[x] = (uint32_t)P->files.size();
        P->files.emplace_back(g, e);
        g = e;
    }
    P->pending.reset(new std::atomic<int64_t>[P->files.size() + 1]);
    for (size_t f = 0; f <= P->files.size(); ++f) P->pending[f].store(0);
    const int nt = std::max(1, threads);
    P->scratch.resize((size_t)nt);
    for (int t = 0; t < nt; ++t) P->threads.emplace_back(&HbFmeaPool::worker, P, t);
    return P;
}

// One block: the HSPs of the query segments that have any, grouped by query segment (segment g owns hsp[q_off[g] ..
// q_off[g+1])).  The arrays stay the caller's and must live until hb_fmea_pool_finish; a block holds whole files and a query
// segment's HSPs come in one block.
void hb_fmea_pool_submit(HbFmeaPool* P, const hb_hsp* hsp, const uint64_t* q_off) {
    std::vector<uint32_t> mine;
    for (uint32_t g = 0; g < P->n_seg; ++g)
        if (q_off[g + 1] > q_off[g]) { P->q_hsp[g] = hsp + q_off[g]; P->q_n[g] = q_off[g + 1] - q_off[g]; mine.push_back(g); }
    if (mine.empty()) return;
    for (uint32_t g : mine) P->pending[P->file_of[g]].fetch_add(1);      // every G holds its file open until its C items are queued
    {
        std::lock_guard<std::mutex> lk(P->mu);
        if (!P->any) { P->any = true; P->t_first = std::chrono::steady_clock::now(); }
        for (uint32_t g : mine) P->push_locked(HbFmeaPool::Task{0, g, 0}, false);
    }
    P->cv.notify_all();
}

// Waits for every task, joins the pool and frees it.  span_ms: first submit -> last task done; busy_ms: task time summed over
// the threads.
void hb_fmea_pool_finish(HbFmeaPool* P, std::vector<hb_interval>& result, double* span_ms, double* busy_ms) {
    {
        std::lock_guard<std::mutex> lk(P->mu);
        P->closing = true;
    }
    P->cv.notify_all();
    for (auto& th : P->threads) th.join();
    result.clear();
    for (uint32_t g = 0; g < P->n_seg; ++g) result.insert(result.end(), P->per[g].begin(), P->per[g].end());
    if (span_ms) *span_ms = P->any ? std::chrono::duration<double, std::milli>(P->t_last - P->t_first).count() : 0.0;
    if (busy_ms) *busy_ms = (double)P->busy_us.load() * 1e-3;
    delete P;
}

// HSPs must be grouped by query segment: query segment g owns hsp[q_off[g] .. q_off[g+1]).
int hb_fmea_run(const hb_hsp* hsp, const uint64_t* q_off, uint32_t n_seg, const hb_segment* segs, int skip_gap,
                int max_len, int threads, std::vector<hb_interval>& result, const uint32_t* file_of_seg) {
    HbFmeaPool* P = hb_fmea_pool_start(n_seg, segs, file_of_seg, skip_gap, max_len, threads);
    hb_fmea_pool_submit(P, hsp, q_off);
    double span = 0.0, busy = 0.0;
    hb_fmea_pool_finish(P, result, &span, &busy);
    if (getenv("HB_DEBUG_TIMING")) fprintf(stderr, "[hb_fmea] %d threads: span %.3f ms, busy %.1f ms summed\n", std::max(1, threads), span, busy);
    return 0;
}

// ================================================================================================
// get_query_copies (reference module/Util.py:6828-7030): chain the BLAST HSPs of ONE query per subject into
// full-length copies, keep per subject cluster the longest chain, order by length, apply the coverage rules.
namespace {

struct CFrag { int32_t qs, qe, ss, se; double ident; };
struct CBest { int32_t qs = -1, qe = -1, len = -1, ss = -1, se = -1, slen = -1, subj = -1; };

void copies_cluster(std::vector<CFrag>& c, int qthr, int sthr, int subj, std::vector<CBest>& out) {
    std::stable_sort(c.begin(), c.end(), [](const CFrag& a, const CFrag& b) { return a.qs != b.qs ? a.qs < b.qs : a.qe < b.qe; });
    const size_t n = c.size();
    std::vector<char> used(n, 0);
    auto mark = [&](size_t i) {
        for (size_t t = 0; t < n; ++t)
            if (c[t].qs == c[i].qs && c[t].qe == c[i].qe && c[t].ss == c[i].ss && c[t].se == c[i].se && c[t].ident == c[i].ident) used[t] = 1;
    };
    CBest best;
    for (size_t i = 0; i < n; ++i) {
        if (used[i]) continue;
        int32_t qs = c[i].qs, qe = c[i].qe, ss = c[i].ss, se = c[i].se;
        int32_t len = qe - qs + 1;                                   // :6919
        mark(i);
        for (size_t j = i + 1; j < n; ++j) {
            if (used[j]) continue;
            const CFrag& f = c[j];
            if (f.qe <= qe) continue;
            const bool fwd = ss < se && f.ss < f.se, rev = ss > se && f.ss > f.se;
            if (fwd && f.se > se) {
                if (f.qs - qe < qthr && f.ss - se < sthr) { qe = f.qe; ss = std::min(ss, f.ss); se = f.se; len = qe - qs; mark(j); }
                else if (f.qs - qe >= qthr) break;
            } else if (rev && f.se < se) {
                if (f.qs - qe < qthr && se - f.ss < sthr) { qe = f.qe; ss = std::max(ss, f.ss); se = f.se; len = qe - qs; mark(j); }
                else if (f.qs - qe >= qthr) break;
            }
        }
        if (len > best.len) { best.qs = qs; best.qe = qe; best.len = len; best.ss = ss; best.se = se; best.slen = std::abs(se - ss) + 1; best.subj = subj; }
    }
    if (best.len != -1) out.push_back(best);
}

}  // namespace

int hb_query_copies_run(const hb_hsp* h, const double* ident, size_t n, int query_len, const int32_t* subj_len,
                        double query_coverage, double subject_coverage, int qthr, int sthr, int max_copy_num,
                        std::vector<hb_query_copy>& result) {
    std::vector<int32_t> order, slot;
    std::vector<std::vector<CFrag>> per;
    for (size_t i = 0; i < n; ++i) {
        const hb_hsp& x = h[i];
        if ((size_t)x.subj >= slot.size()) slot.resize((size_t)x.subj + 1, -1);
        if (slot[x.subj] < 0) { slot[x.subj] = (int32_t)order.size(); order.push_back(x.subj); per.emplace_back(); }
        per[slot[x.subj]].push_back(CFrag{x.qs, x.qe, x.ss, x.se, ident[i]});
    }
    std::vector<CBest> longest;
    std::vector<CFrag> cl;
    for (size_t oi = 0; oi < order.size(); ++oi) {
        std::vector<CFrag> fw, rv;
        for (const CFrag& f : per[oi]) (f.ss > f.se ? rv : fw).push_back(f);
        std::stable_sort(fw.begin(), fw.end(), [](const CFrag& a, const CFrag& b) { return a.ss != b.ss ? a.ss < b.ss : a.se < b.se; });
        std::stable_sort(rv.begin(), rv.end(), [](const CFrag& a, const CFrag& b) { return a.ss != b.ss ? a.ss > b.ss : a.se > b.se; });
        for (int pass = 0; pass < 2; ++pass) {
            const std::vector<CFrag>& p = pass ? rv : fw;
            cl.clear();
            for (size_t k = 0; k < p.size(); ++k) {
                bool closed = false;
                if (k > 0)
                    for (size_t e = cl.size(); e-- > 0;) {
                        const int gap = pass ? (cl[e].se - p[k].ss) : (p[k].ss - cl[e].se);
                        if (gap < sthr && p[k].qe > cl[e].qe) { closed = true; break; }
                    }
                if (k > 0 && !closed) { copies_cluster(cl, qthr, sthr, order[oi], longest); cl.clear(); }
                cl.push_back(p[k]);
            }
            if (!cl.empty()) copies_cluster(cl, qthr, sthr, order[oi], longest);
        }
    }
    std::stable_sort(longest.begin(), longest.end(), [](const CBest& a, const CBest& b) { return a.len > b.len; });   // :6997
    result.clear();
    for (const CBest& q : longest) {
        if ((int)result.size() > max_copy_num) break;                                                               // :7002
        int32_t a = q.ss, b = q.se, strand = 0;
        if (a > b) { std::swap(a, b); strand = 1; }
        bool ok = (double)q.len / (double)query_len >= query_coverage;
        if (subject_coverage > 0) ok = ok && ((double)q.slen / (double)subj_len[q.subj] >= subject_coverage);
        if (!ok) continue;
        bool dup = false;
        for (const hb_query_copy& o : result) if (o.subj == q.subj && o.start == a && o.end == b) { dup = true; break; }
        if (!dup) result.push_back(hb_query_copy{q.subj, a, b, q.len, strand});
    }
    return 0;
}

// ------------------------------------------------------------------------------------------------
// get_full_length_copies_from_blastn_v1 (reference module/Util.py:5907-6128) for one query, as the known-TE genome
// masking drives it (mask_genome_intactTE :6389-6431 -> generate_full_length_out_v1 :6288-6316, search_struct = False).
// Host code for the same reason as FMEA: value-keyed visited sets and data-dependent early exits over a few dozen HSPs
// per (query, subject) -- the GPU's share of the masking is the HSP generation (hb_query_hsps) and the N-fill.
namespace {
struct MFrag { int32_t qs, qe, ss, se; };
inline bool same_frag(const MFrag& a, const MFrag& b) { return a.qs == b.qs && a.qe == b.qe && a.ss == b.ss && a.se == b.se; }

// one cluster of one subject (:6021-6094): greedy extension from every not-yet-visited fragment in query order
void full_length_cluster(std::vector<MFrag>& cl, double skip_gap, double thr, int query_len, int32_t subj,
                         std::vector<hb_query_copy>& result) {
    std::stable_sort(cl.begin(), cl.end(), [](const MFrag& a, const MFrag& b) { return a.qs != b.qs ? a.qs < b.qs : a.qe < b.qe; });
    std::vector<MFrag> visited;                       // the reference keys its visited dict by the tuple's value
    auto seen = [&](const MFrag& f) { for (const MFrag& v : visited) if (same_frag(v, f)) return true; return false; };
    for (size_t i = 0; i < cl.size(); ++i) {
        if (seen(cl[i])) continue;
        int32_t pqs = cl[i].qs, pqe = cl[i].qe, pss = cl[i].ss, pse = cl[i].se;
        int32_t longest = std::abs(pqe - pqs);
        visited.push_back(cl[i]);
        for (size_t j = i + 1; j < cl.size(); ++j) {
            const MFrag& c = cl[j];
            if (seen(c)) continue;
            if (c.qe <= pqe) continue;
            const bool fwd = pss < pse && c.ss < c.se, rev = pss > pse && c.ss > c.se;
            if (!(fwd ? c.se > pse : rev ? c.se < pse : false)) continue;
            const double sgap = fwd ? (double)(c.ss - pse) : (double)(pse - c.ss);
            if ((double)(c.qs - pqe) < skip_gap && sgap < skip_gap) {
                pqe = c.qe;
                pss = fwd ? std::min(pss, c.ss) : std::max(pss, c.ss);
                pse = c.se;
                longest = pqe - pqs;
                visited.push_back(c);
            } else if ((double)(c.qs - pqe) >= skip_gap) break;
        }
        if ((double)longest < thr * (double)query_len) continue;                                   // :6098
        if (!((double)longest / (double)query_len >= thr)) continue;                               // :6123-6124
        const int32_t strand = pss > pse ? 1 : 0;
        const int32_t start = (strand ? pse : pss) - 1, end = strand ? pss : pse;                  // :6105-6112
        bool dup = false;
        for (hb_query_copy& o : result) if (o.subj == subj && o.start == start && o.end == end) { o.strand = strand; dup = true; break; }
        if (!dup) result.push_back(hb_query_copy{subj, start, end, longest, strand});
    }
}
}  // namespace

int hb_full_length_copies_run(const hb_hsp* h, size_t n, int query_len, double full_length_threshold,
                              std::vector<hb_query_copy>& result) {
    const double skip_gap = (double)query_len * full_length_threshold;                             // :5947
    std::vector<int32_t> order, slot;
    std::vector<std::vector<MFrag>> per;
    for (size_t i = 0; i < n; ++i) {
        const hb_hsp& x = h[i];
        if ((size_t)x.subj >= slot.size()) slot.resize((size_t)x.subj + 1, -1);
        if (slot[x.subj] < 0) { slot[x.subj] = (int32_t)order.size(); order.push_back(x.subj); per.emplace_back(); }
        per[slot[x.subj]].push_back(MFrag{x.qs, x.qe, x.ss, x.se});
    }
    result.clear();
    std::vector<MFrag> cl;
    for (size_t oi = 0; oi < order.size(); ++oi) {
        std::vector<MFrag> fw, rv;
        for (const MFrag& f : per[oi]) (f.ss > f.se ? rv : fw).push_back(f);                       // :5960-5964
        std::stable_sort(fw.begin(), fw.end(), [](const MFrag& a, const MFrag& b) { return a.ss != b.ss ? a.ss < b.ss : a.se < b.se; });
        std::stable_sort(rv.begin(), rv.end(), [](const MFrag& a, const MFrag& b) { return a.ss != b.ss ? a.ss > b.ss : a.se > b.se; });
        for (int pass = 0; pass < 2; ++pass) {
            const std::vector<MFrag>& p = pass ? rv : fw;
            cl.clear();
            for (size_t k = 0; k < p.size(); ++k) {
                bool closed = false;
                for (size_t e = cl.size(); e-- > 0;) {                                             // :5977-5984
                    const int gap = pass ? (cl[e].se - p[k].ss) : (p[k].ss - cl[e].se);
                    if ((double)gap < skip_gap && p[k].qe > cl[e].qe) { closed = true; break; }
                }
                if (!cl.empty() && !closed) { full_length_cluster(cl, skip_gap, full_length_threshold, query_len, order[oi], result); cl.clear(); }
                cl.push_back(p[k]);
            }
            if (!cl.empty()) full_length_cluster(cl, skip_gap, full_length_threshold, query_len, order[oi], result);
        }
    }
    return 0;
}
