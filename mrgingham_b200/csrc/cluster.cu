// K2 / K2r: exact emulation of the reference's connected-component stage on the SPARSE candidate
// set {(x,y,r) : r > 15} that K1 emits, one CTA per frame.
//
// Why the sparse set is enough (note N2): the reference's traversal
// (find_chessboard_corners.cc:228-267) only ever expands through pixels whose response is > 15;
// pixels with 0 < r <= 15 can be pushed, but when popped they are merely zeroed, which nothing
// observes. So every accepted pixel, every rejected-by-ratio pixel and every seed is a candidate.
//
// What is emulated literally:
//   * raster-order seeding over [8,w-8) x [8,h-8)                         (:332-335)
//   * LIFO traversal, neighbours pushed +x,-x,+y,-y  => visited -y,+y,-x,+x (:252-255)
//   * membership test at POP time against the RUNNING max: r > 15 && r > (max >> 4) (:159-171)
//   * rejected pixels are zeroed and not expanded                         (:243-247)
//   * a neighbour outside [7,w-7) x [7,h-7) poisons the component          (:216-221)
//   * accept iff !poisoned && N >= 2 && max > 120 && 21x21 variance > 400  (:193-209, :50-88)
//   * centroid = sum(r*x)/sum(r) in double, rescale, x1000 and truncate   (:262-263,:278-279,:350-351)
// The explicit stack of the reference (which may hold duplicates) is replaced by a stackless
// depth-first walk with a parent link and a next-direction counter per candidate: a LIFO stack
// whose stale duplicates pop as no-ops visits pixels in exactly the order of a recursive DFS
// that tries -y,+y,-x,+x and skips pixels that are dead by the time their turn comes.
#include "kernels.cuh"

namespace mrgb200
{

constexpr uint32_t kEmpty = 0xFFFFFFFFu;
constexpr uint32_t kRoot  = 0x1FFFFFFFu;
constexpr int      kClusterThreads = 256;
constexpr int      kRefineMaxPoints = 1024;   // refine: points per frame the parallel path groups (more: one thread, as before)
constexpr int      kRefineStack = 192;         // refine: stack of one point's region flood

struct Component
{
    unsigned long long swx, swy, sw;
    int  n, peak, peak_x, peak_y;
    bool poisoned;
};

// Accepted-so-far component, parked in global scratch between the sequential walk and the
// parallel variance gate
struct ComponentRecord
{
    unsigned long long swx, swy, sw;
    uint32_t peak_xy;   // y << 16 | x
    int32_t  tag;       // find: unused; refine: point index. Set to -1 when the variance gate fails
};

__device__ __forceinline__ uint32_t hash_slot(uint32_t key, int bits) { return (key * 2654435761u) >> (32 - bits); }

__device__ __forceinline__ int table_lookup(const cand_t* cand, const uint32_t* table, int bits, uint32_t key)
{
    const uint32_t mask = (1u << bits) - 1;
    uint32_t slot = hash_slot(key, bits);
    for (;;)
    {
        const uint32_t i = table[slot];
        if (i == kEmpty) return -1;
        // (only the key's bytes are read: in the refinement kernel other threads clear the RESPONSE bytes of candidates
        // of their own regions while this probe passes over them)
        const uint16_t* hw = reinterpret_cast<const uint16_t*>(&cand[i]);
        if (((uint32_t)hw[1] | ((uint32_t)hw[2] << 16)) == key) return (int)i;
        slot = (slot + 1) & mask;
    }
}

// Grows one component from `nroots` start pixels (candidate indices, in the order they are to be
// VISITED, i.e. reverse push order). Single thread.
__device__ void grow_component(Component& c, cand_t* cand, const uint32_t* table, int bits, uint32_t* dfs,
                               const int* roots, int nroots, int w, int h)
{
    c.swx = c.swy = c.sw = 0; c.n = 0; c.peak = 0; c.peak_x = c.peak_y = 0; c.poisoned = false;
    for (int ir = 0; ir < nroots; ir++)
    {
        uint32_t cur = (uint32_t)roots[ir], parent = kRoot;
        for (;;)
        {
            // ---- "pop" cur
            const cand_t cc = cand[cur];
            const int r = cand_r(cc);
            const bool member = r != 0 && r > (c.peak >> 4);   // alive => r > 15 already
            if (r != 0) cand[cur] = cc & ~0xFFFFull;           // member or not, it is zeroed
            if (member)
            {
                const int x = cand_x(cc), y = cand_y(cc);
                if (r > c.peak) { c.peak = r; c.peak_x = x; c.peak_y = y; }
                c.swx += (unsigned long long)(r * x);
                c.swy += (unsigned long long)(r * y);
                c.sw  += (unsigned long long)r;
                c.n++;
                if (x + 1 >= w - kMargin || x - 1 < kMargin || y + 1 >= h - kMargin || y - 1 < kMargin)
                    c.poisoned = true;
                dfs[cur] = parent << 3;
            }
            else
                cur = parent;

            // ---- advance to the next pixel to pop, climbing back up as directions run out
            bool found = false;
            while (cur != kRoot)
            {
                const uint32_t st = dfs[cur];
                const int d = (int)(st & 7);
                if (d == 4) { cur = st >> 3; continue; }
                dfs[cur] = st + 1;
                const cand_t pc = cand[cur];
                int nx = cand_x(pc), ny = cand_y(pc);
                if      (d == 0) ny -= 1;
                else if (d == 1) ny += 1;
                else if (d == 2) nx -= 1;
                else             nx += 1;
                if (nx < kMargin || nx >= w - kMargin || ny < kMargin || ny >= h - kMargin) continue;
                const int q = table_lookup(cand, table, bits, ((uint32_t)ny << 16) | (uint32_t)nx);
                if (q < 0 || cand_r(cand[q]) == 0) continue;   // never pushed, or a stale stack entry
                parent = cur; cur = (uint32_t)q; found = true;
                break;
            }
            if (!found) break;
        }
    }
}

// grow_component() as ONE flat loop: every iteration is either a "pop" (membership test of `cur`) or one step of
// the search for the next pixel (one direction of one node). The refinement kernel runs one of these per lane;
// lanes whose walks are at different depths still execute the same loop body, so a warp is not serialised the
// way it is by the nested, data-dependent loops above. Same visits in the same order.
__device__ void grow_component_flat(Component& c, cand_t* cand, const uint32_t* table, int bits, uint32_t* dfs,
                                    const int* roots, int nroots, int w, int h)
{
    c.swx = c.swy = c.sw = 0; c.n = 0; c.peak = 0; c.peak_x = c.peak_y = 0; c.poisoned = false;
    int ir = 0;
    uint32_t cur = kRoot, parent = kRoot;
    bool pop = false;
    for (;;)
    {
        if (cur == kRoot && !pop)
        {
            if (ir == nroots) break;
            cur = (uint32_t)roots[ir++]; parent = kRoot; pop = true;
        }
        if (pop)
        {
            const cand_t cc = cand[cur];
            const int r = cand_r(cc);
            const bool member = r != 0 && r > (c.peak >> 4);   // alive => r > 15 already
            if (r != 0) *reinterpret_cast<uint16_t*>(&cand[cur]) = 0;   // member or not, it is zeroed (the response bytes only)
            if (member)
            {
                const int x = cand_x(cc), y = cand_y(cc);
                if (r > c.peak) { c.peak = r; c.peak_x = x; c.peak_y = y; }
                c.swx += (unsigned long long)(r * x);
                c.swy += (unsigned long long)(r * y);
                c.sw  += (unsigned long long)r;
                c.n++;
                if (x + 1 >= w - kMargin || x - 1 < kMargin || y + 1 >= h - kMargin || y - 1 < kMargin)
                    c.poisoned = true;
                dfs[cur] = parent << 3;
            }
            else
                cur = parent;
            pop = false;
            continue;
        }
        // one step of the advance: the next direction of `cur`, or back to its parent
        const uint32_t st = dfs[cur];
        const int d = (int)(st & 7);
        if (d == 4) { cur = st >> 3; continue; }
        dfs[cur] = st + 1;
        const cand_t pc = cand[cur];
        int nx = cand_x(pc), ny = cand_y(pc);
        if      (d == 0) ny -= 1;
        else if (d == 1) ny += 1;
        else if (d == 2) nx -= 1;
        else             nx += 1;
        if (nx < kMargin || nx >= w - kMargin || ny < kMargin || ny >= h - kMargin) continue;
        const int q = table_lookup(cand, table, bits, ((uint32_t)ny << 16) | (uint32_t)nx);
        if (q < 0 || cand_r(cand[q]) == 0) continue;           // never pushed, or a stale stack entry
        parent = cur; cur = (uint32_t)q; pop = true;
    }
}

// The same traversal as grow_component(), for a region that has been collected into a thread-local
// graph: node i has raster key lkey[i] (y<<16|x), live response lr[i] (0 once visited) and the local
// indices of its -y,+y,-x,+x neighbours in nb[i][0..3] (-1 = no candidate there). No hash lookups.
constexpr int kRegionMax = 48;
struct LocalRegion
{
    uint32_t lkey[kRegionMax];
    int8_t   nb[kRegionMax][4];       // read as one 32-bit word per node: keep 4-byte aligned (right after lkey)
    uint16_t lr[kRegionMax];
    int8_t   dfs_parent[kRegionMax];
    uint8_t  dfs_dir[kRegionMax];
};

__device__ void grow_component_local(Component& c, LocalRegion& g, int seed, int w, int h)
{
    // Stackless depth-first replay of follow_connected_component() (find_chessboard_corners.cc:228-267) on
    // the region's local graph, as ONE flat loop: an iteration is either a "pop" (membership test of `cur`) or one
    // direction of the node being expanded. Lanes replaying different regions then execute the same loop body
    // whatever the depth of their walks; the nested loops this replaces left 3-7 active lanes per warp.
    // The node being expanded keeps its four neighbour indices (one 32-bit load) and its next direction in
    // registers; per-node state goes to the local arrays only when the walk descends to a child.
    c.swx = c.swy = c.sw = 0; c.n = 0; c.peak = 0; c.peak_x = c.peak_y = 0; c.poisoned = false;
    const uint32_t* nbw_of = reinterpret_cast<const uint32_t*>(&g.nb[0][0]);
    int cur = seed, parent = -1, d = 0;
    uint32_t nbw = 0;
    bool pop = true;
    for (;;)
    {
        if (pop)
        {
            const int r = g.lr[cur];
            const bool member = r != 0 && r > (c.peak >> 4);
            g.lr[cur] = 0;
            if (member)
            {
                const int x = (int)(g.lkey[cur] & 0xFFFF), y = (int)(g.lkey[cur] >> 16);
                if (r > c.peak) { c.peak = r; c.peak_x = x; c.peak_y = y; }
                c.swx += (unsigned long long)(r * x);
                c.swy += (unsigned long long)(r * y);
                c.sw  += (unsigned long long)r;
                c.n++;
                if (x + 1 >= w - kMargin || x - 1 < kMargin || y + 1 >= h - kMargin || y - 1 < kMargin)
                    c.poisoned = true;
                g.dfs_parent[cur] = (int8_t)parent;
                d = 0; nbw = nbw_of[cur];
            }
            else
            {
                cur = parent;
                if (cur < 0) break;
                d = g.dfs_dir[cur]; nbw = nbw_of[cur];
            }
            pop = false;
            continue;
        }
        if (d == 4)
        {
            cur = g.dfs_parent[cur];
            if (cur < 0) break;
            d = g.dfs_dir[cur]; nbw = nbw_of[cur];
            continue;
        }
        const int q = (int)(int8_t)(nbw >> (8 * d));
        d++;
        if (q < 0 || g.lr[q] == 0) continue;
        g.dfs_dir[cur] = (uint8_t)d;          // where to resume when the walk comes back to this node
        parent = cur; cur = q; pop = true;
    }
}

// 21x21 variance gate around (x,y) of the level image, one warp (find_chessboard_corners.cc:50-88)
__device__ bool variance_gate_warp(const uint8_t* img, int pitch, int w, int h, int x, int y, int lane)
{
    if (x - kVarWindowR < 0 || x + kVarWindowR >= w || y - kVarWindowR < 0 || y + kVarWindowR >= h)
        return false;
    constexpr int D = 2*kVarWindowR + 1, NPIX = D*D;
    int vals[(NPIX + 31) / 32];
    int sum = 0;
#pragma unroll
    for (int k = 0; k < (NPIX + 31) / 32; k++)
    {
        const int i = lane + 32*k;
        int v = 0;
        if (i < NPIX) v = img[(size_t)(y - kVarWindowR + i / D) * pitch + (x - kVarWindowR + i % D)];
        vals[k] = v; sum += v;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    const int mean = sum / NPIX;
    int ssd = 0;
#pragma unroll
    for (int k = 0; k < (NPIX + 31) / 32; k++)
        if (lane + 32*k < NPIX) { const int e = vals[k] - mean; ssd += e*e; }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) ssd += __shfl_xor_sync(0xffffffffu, ssd, o);
    return ssd / NPIX > kVarMin;
}

__device__ __forceinline__ double rescale_coord(double p, double scale)
{
    // (p + 0.5)*scale - 0.5 with every operation rounded separately (no FMA contraction), as the
    // reference's x86-64 build does (find_chessboard_corners.cc:278-279)
    return __dadd_rn(__dmul_rn(__dadd_rn(p, 0.5), scale), -0.5);
}

// Everything up to "candidates sorted + hash table built", shared by find and refine.
// Returns false (after writing the overflow marker) if the frame's list overflowed.
struct FrameWork
{
    cand_t*   cand;
    uint32_t* table;
    uint32_t* dfs;
    int       n, bits, P;
    bool      in_smem;
};

// sort_mode: 0 = never sort (caller orders what it needs itself; only valid for the shared-memory
// path), 1 = always sort into raster order, 2 = sort only lists that live in global scratch
__device__ bool prepare_frame(FrameWork& fw, int f, int cap, cand_t* cand_all, const uint32_t* counts,
                              uint32_t* scratch_table, uint32_t* scratch_dfs, uint8_t* smem, int sort_mode, int scap)
{
    const int tid = threadIdx.x;
    const uint32_t total = counts[f];
    if (total > (uint32_t)cap) return false;
    const int n = (int)total;
    int P = 1, lg = 0;
    while (P < n) { P <<= 1; lg++; }
    cand_t* gcand = cand_all + (size_t)f * cap;
    const bool in_smem = P <= scap;          // scap = candidates the launch's shared memory holds (ClusterParams::smem_cands)
    cand_t*   cand  = in_smem ? (cand_t*)smem : gcand;
    uint32_t* table = in_smem ? (uint32_t*)(smem + sizeof(cand_t) * scap) : scratch_table + (size_t)f * 2 * cap;
    uint32_t* dfs   = in_smem ? (uint32_t*)(smem + (sizeof(cand_t) + 2*sizeof(uint32_t)) * scap) : scratch_dfs + (size_t)f * cap;

    for (int i = tid; i < P; i += kClusterThreads)
        cand[i] = i < n ? gcand[i] : ~0ull;
    __syncthreads();

    // bitonic sort: ascending word order == raster order
    if (sort_mode == 1 || (sort_mode == 2 && !in_smem))
    for (int k = 2; k <= P; k <<= 1)
        for (int j = k >> 1; j > 0; j >>= 1)
        {
            for (int i = tid; i < P; i += kClusterThreads)
            {
                const int l = i ^ j;
                if (l > i)
                {
                    const cand_t a = cand[i], b = cand[l];
                    const bool up = (i & k) == 0;
                    if ((a > b) == up) { cand[i] = b; cand[l] = a; }
                }
            }
            __syncthreads();
        }

    const int bits = lg + 1;
    for (int i = tid; i < (1 << bits); i += kClusterThreads) table[i] = kEmpty;
    __syncthreads();
    const uint32_t mask = (1u << bits) - 1;
    for (int i = tid; i < n; i += kClusterThreads)
    {
        uint32_t slot = hash_slot(cand_key(cand[i]), bits);
        while (atomicCAS(&table[slot], kEmpty, (uint32_t)i) != kEmpty) slot = (slot + 1) & mask;
    }
    __syncthreads();
    fw.cand = cand; fw.table = table; fw.dfs = dfs; fw.n = n; fw.bits = bits; fw.P = P; fw.in_smem = in_smem;
    return true;
}

// ------------------------------------------------------------------------------------------------
// K2: find branch
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kClusterThreads)
cluster_find_kernel(FrameSet fs, ClusterParams p, cand_t* cand_all, const uint32_t* counts,
                    uint32_t* scratch_table, uint32_t* scratch_dfs,
                    int32_t* xy_int, double* xy_dbl, int32_t* out_counts)
{
    extern __shared__ __align__(16) uint8_t smem[];
    __shared__ int s_nrec, s_nout, s_nstart;
    const int f = blockIdx.x, tid = threadIdx.x;
    FrameWork fw;
    if (!prepare_frame(fw, f, p.cand_capacity, cand_all, counts, scratch_table, scratch_dfs, smem, 2, p.smem_cands))
    {
        if (tid == 0) out_counts[f] = -1;
        return;
    }
    const int w = fs.w, h = fs.h;
    const int record_cap = p.record_capacity;
    ComponentRecord* rec = (ComponentRecord*)p.records + (size_t)f * record_cap;

    if (tid == 0) s_nrec = 0;
    __syncthreads();

    // One seed's component, grown by the calling thread; parks it if it passes the cheap tests.
    auto try_seed = [&](int s)
    {
        const cand_t sc = fw.cand[s];
        if (cand_r(sc) == 0) return;
        const int x = cand_x(sc), y = cand_y(sc);
        if (x < kMargin + 1 || x >= w - kMargin - 1 || y < kMargin + 1 || y >= h - kMargin - 1) return;
        Component c;
        grow_component(c, fw.cand, fw.table, fw.bits, fw.dfs, &s, 1, w, h);
        if (c.poisoned || c.n < kComponentMinN || c.peak <= kPeakMin) return;
        const int slot = atomicAdd(&s_nrec, 1);
        if (slot < record_cap)
        {
            ComponentRecord r;
            r.swx = c.swx; r.swy = c.swy; r.sw = c.sw;
            r.peak_xy = ((uint32_t)c.peak_y << 16) | (uint32_t)c.peak_x;
            r.tag = (int32_t)cand_key(sc);          // raster position of the seed = output order
            rec[slot] = r;
        }
    };

    bool sequential = !fw.in_smem;
    if (fw.in_smem)
    {
        // The reference's scan is sequential only WITHIN a 4-connected region of the candidate set
        // (nothing a component does reaches outside its region), so regions are replayed in
        // parallel, one thread each. A region is found from its raster-first pixel: every candidate
        // with no candidate above it and none to its left ("starter") floods its region through the
        // hash table and gives up as soon as it meets a pixel that precedes it in raster order, so
        // exactly one starter per region -- its first pixel -- survives with the full member list.
        uint16_t* starters = (uint16_t*)(smem + (sizeof(cand_t) + 3*sizeof(uint32_t)) * p.smem_cands);
        // pointers derived straight from the shared array, so these accesses compile to LDS/STS
        // (the FrameWork pointers are generic: they may also point at global scratch)
        const cand_t*   scand  = (const cand_t*)smem;
        const uint32_t* stable = (const uint32_t*)(smem + sizeof(cand_t) * p.smem_cands);
        const int n = fw.n;
        if (tid == 0) s_nstart = 0;
        __syncthreads();
        for (int i = tid; i < n; i += kClusterThreads)
        {
            const uint32_t key = cand_key(scand[i]);
            if (table_lookup(scand, stable, fw.bits, key - 0x10000u) < 0 &&
                table_lookup(scand, stable, fw.bits, key - 1u) < 0)
                starters[atomicAdd(&s_nstart, 1)] = (uint16_t)i;
        }
        __syncthreads();
        const int nstart = s_nstart;
        // The floods and replays below diverge lane by lane, so a warp's time grows with the number of its lanes
        // that have a starter: deal the starters to the warps round-robin (lane l of warp w takes starter
        // l * nwarps + w, a bijection on 0..kClusterThreads-1) instead of filling warp after warp.
        constexpr int kWarps = kClusterThreads / 32;
        for (int k = (tid & 31) * kWarps + (tid >> 5); k < nstart && !sequential; k += kClusterThreads)
        {
            const int i0 = starters[k];
            const uint32_t key = cand_key(scand[i0]);
            LocalRegion g;
            int gidx[kRegionMax];     // local index -> candidate index
            int count = 1;
            gidx[0] = i0; g.lkey[0] = key;
            bool first = true;
            for (int head = 0; head < count && first && !sequential; head++)
            {
                const uint32_t mk = g.lkey[head];
                const uint32_t nk[4] = { mk - 0x10000u, mk + 0x10000u, mk - 1u, mk + 1u };
#pragma unroll
                for (int d = 0; d < 4; d++)
                {
                    g.nb[head][d] = -1;
                    if (!first || sequential) continue;
                    // already collected? (several starters of one region flood it concurrently, so
                    // the "seen" state has to be private: scan the short local list)
                    int li = -1;
                    for (int e = 0; e < count; e++) if (g.lkey[e] == nk[d]) li = e;
                    if (li < 0)
                    {
                        const int q = table_lookup(scand, stable, fw.bits, nk[d]);
                        if (q < 0) continue;
                        if (nk[d] < key) { first = false; continue; }     // someone precedes this starter
                        if (count == kRegionMax) { sequential = true; continue; }
                        li = count++; gidx[li] = q; g.lkey[li] = nk[d];
                    }
                    g.nb[head][d] = (int8_t)li;
                }
            }
            if (sequential) break;
            if (!first) continue;     // not this region's first pixel: its owner does the work
            for (int a = 0; a < count; a++) g.lr[a] = (uint16_t)cand_r(scand[gidx[a]]);
            // seeds in raster order (insertion sort of the local indices by key)
            int8_t order[kRegionMax];
            for (int a = 0; a < count; a++)
            {
                int b = a - 1;
                while (b >= 0 && g.lkey[order[b]] > g.lkey[a]) { order[b + 1] = order[b]; b--; }
                order[b + 1] = (int8_t)a;
            }
            for (int a = 0; a < count; a++)
            {
                const int sd = order[a];
                if (g.lr[sd] == 0) continue;
                const int x = (int)(g.lkey[sd] & 0xFFFF), y = (int)(g.lkey[sd] >> 16);
                if (x < kMargin + 1 || x >= w - kMargin - 1 || y < kMargin + 1 || y >= h - kMargin - 1) continue;
                Component c;
                grow_component_local(c, g, sd, w, h);
                if (c.poisoned || c.n < kComponentMinN || c.peak <= kPeakMin) continue;
                const int slot = atomicAdd(&s_nrec, 1);
                if (slot < record_cap)
                {
                    ComponentRecord r;
                    r.swx = c.swx; r.swy = c.swy; r.sw = c.sw;
                    r.peak_xy = ((uint32_t)c.peak_y << 16) | (uint32_t)c.peak_x;
                    r.tag = (int32_t)g.lkey[sd];
                    rec[slot] = r;
                }
            }
        }
        // a region too large for one thread's scratch: redo the whole frame the sequential way
        sequential = __syncthreads_or(sequential) != 0;
        if (sequential)
        {
            if (!prepare_frame(fw, f, p.cand_capacity, cand_all, counts, scratch_table, scratch_dfs, smem, 1, p.smem_cands)) return;
            if (tid == 0) s_nrec = 0;
            __syncthreads();
        }
    }
    if (sequential && tid == 0)
        for (int s = 0; s < fw.n; s++) try_seed(s);     // raster order: the list is sorted on this path
    __syncthreads();
    const int nrec = s_nrec;
    if (nrec > record_cap)
    {
        if (tid == 0) out_counts[f] = -1;
        return;
    }

    // variance gate: one warp per record
    const uint8_t* img = fs.base + (size_t)f * fs.frame_stride;
    for (int k = tid >> 5; k < nrec; k += kClusterThreads / 32)
    {
        const uint32_t pk = rec[k].peak_xy;
        const bool ok = variance_gate_warp(img, fs.pitch, w, h, (int)(pk & 0xFFFF), (int)(pk >> 16), tid & 31);
        if ((tid & 31) == 0 && !ok) rec[k].tag = -1;
    }
    __syncthreads();

    // Output order = raster order of the seeds (find_chessboard_corners.cc:332-351): rank the
    // surviving records by seed index.
    if (tid == 0) s_nout = 0;
    // (the hash table is no longer needed: reuse its shared memory for the tags when they fit)
    const bool tags_in_smem = nrec <= 2 * p.smem_cands;
    int32_t* tags = (int32_t*)(smem + sizeof(cand_t) * p.smem_cands);
    if (tags_in_smem)
        for (int k = tid; k < nrec; k += kClusterThreads) tags[k] = rec[k].tag;
    __syncthreads();
    const double scale = (double)(1 << p.level);
    for (int k = tid; k < nrec; k += kClusterThreads)
    {
        const int tag = rec[k].tag;
        if (tag < 0) continue;
        int rank = 0;
        if (tags_in_smem)
            for (int j = 0; j < nrec; j++) { const int tj = tags[j]; rank += (tj >= 0 && tj < tag) ? 1 : 0; }
        else
            for (int j = 0; j < nrec; j++) { const int tj = rec[j].tag; rank += (tj >= 0 && tj < tag) ? 1 : 0; }
        atomicAdd(&s_nout, 1);
        if (rank < p.max_points)
        {
            const double sw = __ull2double_rn(rec[k].sw);
            const double fx = rescale_coord(__ddiv_rn(__ull2double_rn(rec[k].swx), sw), scale);
            const double fy = rescale_coord(__ddiv_rn(__ull2double_rn(rec[k].swy), sw), scale);
            const size_t o = ((size_t)f * p.max_points + rank) * 2;
            xy_int[o]     = __double2int_rz(__dadd_rn(0.5, __dmul_rn(fx, kFindGridScale)));
            xy_int[o + 1] = __double2int_rz(__dadd_rn(0.5, __dmul_rn(fy, kFindGridScale)));
            if (xy_dbl) { xy_dbl[o] = fx; xy_dbl[o + 1] = fy; }
        }
    }
    __syncthreads();
    if (tid == 0) out_counts[f] = s_nout;
}

// ------------------------------------------------------------------------------------------------
// K2r: refine branch (find_chessboard_corners.cc:356-397). Points are visited in index order on
// one shared, mutable candidate set.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kClusterThreads)
cluster_refine_kernel(FrameSet fs, ClusterParams p, cand_t* cand_all, const uint32_t* counts,
                      uint32_t* scratch_table, uint32_t* scratch_dfs,
                      double* points_xy, signed char* levels, int npoints, int32_t* out_refined)
{
    extern __shared__ __align__(16) uint8_t smem[];
    __shared__ int s_nrec, s_nrefined;
    const int f = blockIdx.x, tid = threadIdx.x;
    if (tid == 0) s_nrefined = 0;
    FrameWork fw;
    if (!prepare_frame(fw, f, p.cand_capacity, cand_all, counts, scratch_table, scratch_dfs, smem, 1, p.smem_cands))
    {
        if (tid == 0) out_refined[f] = -1;
        return;
    }
    const int w = fs.w, h = fs.h;
    const int record_cap = p.record_capacity;
    ComponentRecord* rec = (ComponentRecord*)p.records + (size_t)f * record_cap;
    double*      xy  = points_xy + (size_t)f * npoints * 2;
    signed char* lvl = levels    + (size_t)f * npoints;
    const double scale = (double)(1 << p.level), inv_scale = 1.0 / scale;

    // One point's visit of the shared candidate set: its 3x3 seeds, the component grown from them, a record if the
    // component passes. Points whose seeds lie in different 4-connected regions of the candidate set cannot see
    // each other's visits (nothing a component does leaves its region), so they may run concurrently.
    auto visit_point = [&](int i)
    {
        const int x = __double2int_rz(__dadd_rn(rescale_coord(xy[2*i],     inv_scale), 0.5));
        const int y = __double2int_rz(__dadd_rn(rescale_coord(xy[2*i + 1], inv_scale), 0.5));
        // 3x3 seeds, pushed dx-outer / dy-inner, hence visited in the reverse order
        int roots[9], nroots = 0;
        for (int dx = 1; dx >= -1; dx--)
            for (int dy = 1; dy >= -1; dy--)
            {
                const int u = x + dx, v = y + dy;
                if (u < kMargin || u >= w - kMargin || v < kMargin || v >= h - kMargin) continue; // response is 0 there
                const int q = table_lookup(fw.cand, fw.table, fw.bits, ((uint32_t)v << 16) | (uint32_t)u);
                if (q >= 0 && cand_r(fw.cand[q]) != 0) roots[nroots++] = q;
            }
        Component c;
        grow_component_flat(c, fw.cand, fw.table, fw.bits, fw.dfs, roots, nroots, w, h);
        if (c.poisoned || c.n < kComponentMinN || c.peak <= kPeakMin) return;
        const int k = atomicAdd(&s_nrec, 1);
        if (k < record_cap)
        {
            ComponentRecord r;
            r.swx = c.swx; r.swy = c.swy; r.sw = c.sw;
            r.peak_xy = ((uint32_t)c.peak_y << 16) | (uint32_t)c.peak_x; r.tag = i;
            rec[k] = r;
        }
    };

    // Which points share a region? Every point floods the regions of its seeds over the STATIC candidate graph,
    // lowering a per-candidate mark (kept in the dfs array, which the visits only use afterwards) to its own index
    // and expanding only where it lowered one: when all floods are done every candidate of a region carries the
    // smallest index of the points seeded in it. Points are then united through the marks of their seeds; each
    // group is visited, in index order, by one thread.
    __shared__ int s_fallback;
    __shared__ uint16_t s_parent[kRefineMaxPoints];
    if (tid == 0) { s_nrec = 0; s_fallback = npoints > kRefineMaxPoints || fw.n > 65535; }
    for (int i = tid; i < fw.n; i += kClusterThreads) fw.dfs[i] = 0xFFFFFFFFu;
    __syncthreads();
    if (!s_fallback)
    {
        for (int i = tid; i < npoints; i += kClusterThreads)
        {
            s_parent[i] = (uint16_t)i;
            if (lvl[i] != p.level + 1) continue;
            const int x = __double2int_rz(__dadd_rn(rescale_coord(xy[2*i],     inv_scale), 0.5));
            const int y = __double2int_rz(__dadd_rn(rescale_coord(xy[2*i + 1], inv_scale), 0.5));
            // a candidate is claimed (its mark lowered) when it is PUSHED, so it is on the stack at most once
            uint16_t stk[kRefineStack];
            int sp = 0;
            bool overflow = false;
            auto claim = [&](int q)
            {
                if (atomicMin(&fw.dfs[q], (uint32_t)i) <= (uint32_t)i) return;             // a point with a smaller index owns it (or this one does already)
                if (sp == kRefineStack) { overflow = true; return; }
                stk[sp++] = (uint16_t)q;
            };
            for (int dx = -1; dx <= 1; dx++)
                for (int dy = -1; dy <= 1; dy++)
                {
                    const int u = x + dx, v = y + dy;
                    if (u < kMargin || u >= w - kMargin || v < kMargin || v >= h - kMargin) continue;
                    const int q = table_lookup(fw.cand, fw.table, fw.bits, ((uint32_t)v << 16) | (uint32_t)u);
                    if (q >= 0) claim(q);
                }
            while (sp > 0 && !overflow)
            {
                const cand_t cc = fw.cand[stk[--sp]];
                const int cx = cand_x(cc), cy = cand_y(cc);
#pragma unroll
                for (int d = 0; d < 4; d++)
                {
                    const int nx = cx + (d == 2 ? -1 : d == 3 ? 1 : 0), ny = cy + (d == 0 ? -1 : d == 1 ? 1 : 0);
                    if (nx < kMargin || nx >= w - kMargin || ny < kMargin || ny >= h - kMargin) continue;
                    const int q = table_lookup(fw.cand, fw.table, fw.bits, ((uint32_t)ny << 16) | (uint32_t)nx);
                    if (q >= 0) claim(q);
                }
            }
            if (overflow) s_fallback = 1;
        }
    }
    __syncthreads();
    if (s_fallback)
    {
        // (more points than the group table holds, or a region too large for a flood's stack: one thread, index order)
        if (tid == 0)
            for (int i = 0; i < npoints; i++)
                if (lvl[i] == p.level + 1) visit_point(i);
    }
    else
    {
        // a point joins the owner of its seeds' region (the mark; <= its own index: it flooded them too). Points whose
        // seeds carry different marks bridge regions: those few unions are left to one thread.
        __shared__ int s_nbridge;
        __shared__ uint16_t s_bridge[kRefineMaxPoints];
        if (tid == 0) s_nbridge = 0;
        __syncthreads();
        for (int i = tid; i < npoints; i += kClusterThreads)
        {
            if (lvl[i] != p.level + 1) continue;
            const int x = __double2int_rz(__dadd_rn(rescale_coord(xy[2*i],     inv_scale), 0.5));
            const int y = __double2int_rz(__dadd_rn(rescale_coord(xy[2*i + 1], inv_scale), 0.5));
            uint32_t lo = (uint32_t)i, hi = 0; bool any = false;
            for (int dx = -1; dx <= 1; dx++)
                for (int dy = -1; dy <= 1; dy++)
                {
                    const int u = x + dx, v = y + dy;
                    if (u < kMargin || u >= w - kMargin || v < kMargin || v >= h - kMargin) continue;
                    const int q = table_lookup(fw.cand, fw.table, fw.bits, ((uint32_t)v << 16) | (uint32_t)u);
                    if (q < 0) continue;
                    const uint32_t m = fw.dfs[q];
                    lo = min(lo, m); hi = max(hi, m); any = true;
                }
            s_parent[i] = (uint16_t)lo;
            if (any && hi != lo) s_bridge[atomicAdd(&s_nbridge, 1)] = (uint16_t)i;
        }
        __syncthreads();
        if (tid == 0 && s_nbridge > 0)
        {
            auto find = [&](int a) { while (s_parent[a] != a) { s_parent[a] = s_parent[s_parent[a]]; a = s_parent[a]; } return a; };
            for (int b = 0; b < s_nbridge; b++)
            {
                const int i = s_bridge[b];
                const int x = __double2int_rz(__dadd_rn(rescale_coord(xy[2*i],     inv_scale), 0.5));
                const int y = __double2int_rz(__dadd_rn(rescale_coord(xy[2*i + 1], inv_scale), 0.5));
                for (int dx = -1; dx <= 1; dx++)
                    for (int dy = -1; dy <= 1; dy++)
                    {
                        const int u = x + dx, v = y + dy;
                        if (u < kMargin || u >= w - kMargin || v < kMargin || v >= h - kMargin) continue;
                        const int q = table_lookup(fw.cand, fw.table, fw.bits, ((uint32_t)v << 16) | (uint32_t)u);
                        if (q < 0) continue;
                        const int ra = find(i), rb = find((int)fw.dfs[q]);
                        if (ra != rb) s_parent[max(ra, rb)] = (uint16_t)min(ra, rb);
                    }
            }
        }
        __syncthreads();
        // every point's group = the root of its chain (chains only ever point to smaller indices)
        int my_root[(kRefineMaxPoints + kClusterThreads - 1) / kClusterThreads];
#pragma unroll
        for (int k = 0; k < (kRefineMaxPoints + kClusterThreads - 1) / kClusterThreads; k++)
        {
            const int i = tid + k * kClusterThreads;
            int a = i < npoints ? i : 0;
            while (s_parent[a] != a) a = s_parent[a];
            my_root[k] = a;
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < (kRefineMaxPoints + kClusterThreads - 1) / kClusterThreads; k++)
        {
            const int i = tid + k * kClusterThreads;
            if (i < npoints) s_parent[i] = (uint16_t)my_root[k];
        }
        __syncthreads();
        // (points dealt to the warps round-robin: fewer diverging walks per warp)
        for (int i0 = 0; i0 < npoints; i0 += kClusterThreads)
        {
            const int i = i0 + (tid & 31) * (kClusterThreads / 32) + (tid >> 5);
            if (i >= npoints || s_parent[i] != i) continue;    // the group's first point leads it
            for (int j = i; j < npoints; j++)
                if (s_parent[j] == i && lvl[j] == p.level + 1) visit_point(j);
        }
    }
    __syncthreads();
    const int nrec = s_nrec;
    if (nrec > record_cap)
    {
        if (tid == 0) out_refined[f] = -1;
        return;
    }
    const uint8_t* img = fs.base + (size_t)f * fs.frame_stride;
    for (int k = tid >> 5; k < nrec; k += kClusterThreads / 32)
    {
        const uint32_t pk = rec[k].peak_xy;
        const bool ok = variance_gate_warp(img, fs.pitch, w, h, (int)(pk & 0xFFFF), (int)(pk >> 16), tid & 31);
        if ((tid & 31) == 0 && ok)
        {
            const int i = rec[k].tag;
            const double sw = __ull2double_rn(rec[k].sw);
            xy[2*i]     = rescale_coord(__ddiv_rn(__ull2double_rn(rec[k].swx), sw), scale);
            xy[2*i + 1] = rescale_coord(__ddiv_rn(__ull2double_rn(rec[k].swy), sw), scale);
            lvl[i] = (signed char)p.level;
            atomicAdd(&s_nrefined, 1);
        }
    }
    __syncthreads();
    if (tid == 0) out_refined[f] = s_nrefined;
}

static size_t cluster_smem_bytes(int scap) { return (sizeof(cand_t) + 3*sizeof(uint32_t) + sizeof(uint16_t)) * (size_t)scap; }
static int checked_smem_cands(const ClusterParams& p)
{
    int c = p.smem_cands;
    if (c != 1024 && c != 2048 && c != 4096) c = kClusterSmemCands;
    return c;
}

size_t cluster_record_bytes() { return sizeof(ComponentRecord); }

cudaError_t launch_cluster_find(const FrameSet& fs, const ClusterParams& p_in,
                                cand_t* cand, const uint32_t* counts,
                                uint32_t* scratch_table, uint32_t* scratch_dfs,
                                int32_t* xy_int, double* xy_dbl, int32_t* out_counts,
                                cudaStream_t stream)
{
    if (fs.nframes <= 0) return cudaSuccess;
    ClusterParams p = p_in;
    p.smem_cands = checked_smem_cands(p_in);
    const size_t smem = cluster_smem_bytes(p.smem_cands);
    if (smem > 48 * 1024)
    {
        // per device, idempotent and cheap: set on every launch rather than tracking devices
        cudaError_t e = cudaFuncSetAttribute(cluster_find_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
    }
    cluster_find_kernel<<<fs.nframes, kClusterThreads, smem, stream>>>(fs, p, cand, counts, scratch_table, scratch_dfs,
                                                                       xy_int, xy_dbl, out_counts);
    return cudaGetLastError();
}

cudaError_t launch_cluster_refine(const FrameSet& fs, const ClusterParams& p_in,
                                  cand_t* cand, const uint32_t* counts,
                                  uint32_t* scratch_table, uint32_t* scratch_dfs,
                                  double* points_xy, signed char* levels, int npoints,
                                  int32_t* out_refined, cudaStream_t stream)
{
    if (fs.nframes <= 0) return cudaSuccess;
    ClusterParams p = p_in;
    p.smem_cands = checked_smem_cands(p_in);
    const size_t smem = cluster_smem_bytes(p.smem_cands);
    if (smem + 8 * 1024 > 48 * 1024)       // (the kernel's static arrays count against the 48 KB default too)
    {
        // per device, idempotent and cheap: set on every launch rather than tracking devices
        cudaError_t e = cudaFuncSetAttribute(cluster_refine_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
    }
    cluster_refine_kernel<<<fs.nframes, kClusterThreads, smem, stream>>>(fs, p, cand, counts, scratch_table, scratch_dfs,
                                                                         points_xy, levels, npoints, out_refined);
    return cudaGetLastError();
}

}
