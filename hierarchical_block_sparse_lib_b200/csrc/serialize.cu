// serialize.cu -- the reference's flat wire format (get_size H:1124, write_to_buffer H:1159, assign_from_buffer H:1348):
// the form in which the Chunks-and-Tasks runtime ships a leaf matrix.  Byte-compatible with the reference:
//   node := int nRows, nCols (virtual size at this level), int nRows_orig, nCols_orig, int blocksize,
//           Treal frob_norm_squared_internal, size_t n_block_multiplies, size_t child_size[4],
//           then either blocksize^2 Treal (leaf, column-major) or the existing children 0..3 (same layout)
// The quadtree only exists on the wire: it is rebuilt from / flattened into the Morton block table here (host code; the
// tiles cross PCIe once).  Inner-node norms on the wire are the hierarchical sums of the cached leaf norms (what
// update_internal_info leaves behind, H:3918-3923); the root carries the matrix' cached root norm and multiply counter.
// Known divergence: the flat table keeps ONE n_block_multiplies per matrix, so non-root nodes are written with counter 0; the
// reference's product results carry per-node counters on inner nodes of C (H:2186-2194 sets the root only, but add() sums
// children's counters, H:1719).  Root counter, structure, norms and values are byte-identical (tests/golden/wire_format_*).
#include "matrix.cuh"

namespace hbsm_b200 {

namespace {

struct Flat {
    int b = 0, M = 0, N = 0, depth = 0;
    size_t esize = 8, tile_bytes = 0;
    std::vector<uint64_t> keys;
    std::vector<char> norms, tiles;
    double root_norm = 0.0;
    size_t n_mults = 0;
};

size_t header_bytes(size_t esize) { return 5 * sizeof(int) + esize + sizeof(size_t) + 4 * sizeof(size_t); }

// size of the subtree holding leaves [lo,hi) whose root sits `level` levels above the leaves
size_t subtree_size(const Flat& f, size_t lo, size_t hi, int level) {
    if (level == 0) return header_bytes(f.esize) + f.tile_bytes;
    size_t total = header_bytes(f.esize);
    size_t p = lo;
    for (int q = 0; q < 4; ++q) {
        size_t e = p;
        while (e < hi && ((f.keys[e] >> (2 * (level - 1))) & 3u) == (uint64_t)q) ++e;
        if (e > p) total += subtree_size(f, p, e, level - 1);
        p = e;
    }
    return total;
}

double norm_at(const Flat& f, size_t i) {
    return f.esize == 8 ? reinterpret_cast<const double*>(f.norms.data())[i] : (double)reinterpret_cast<const float*>(f.norms.data())[i];
}

// cached norm of a subtree = sum of the children's cached norms in child order, in Treal (H:3918-3923)
double subtree_norm(const Flat& f, size_t lo, size_t hi, int level) {
    if (level == 0) return norm_at(f, lo);
    double sd = 0.0; float sf = 0.0f;
    size_t p = lo;
    for (int q = 0; q < 4; ++q) {
        size_t e = p;
        while (e < hi && ((f.keys[e] >> (2 * (level - 1))) & 3u) == (uint64_t)q) ++e;
        if (e > p) { double c = subtree_norm(f, p, e, level - 1); sd += c; sf += (float)c; }
        p = e;
    }
    return f.esize == 8 ? sd : (double)sf;
}

template <typename V> void put(char*& p, const V& v) { memcpy(p, &v, sizeof(V)); p += sizeof(V); }
template <typename V> V take(const char*& p) { V v; memcpy(&v, p, sizeof(V)); p += sizeof(V); return v; }

void put_real(char*& p, size_t esize, double v) {
    if (esize == 8) put(p, v);
    else put(p, (float)v);
}

void write_node(const Flat& f, size_t lo, size_t hi, int level, bool is_root, char*& p) {
    const int vs = f.b << level;
    put(p, vs); put(p, vs);
    put(p, is_root ? f.M : vs); put(p, is_root ? f.N : vs);   // children are resized to their virtual size (H:791-833)
    put(p, f.b);
    put_real(p, f.esize, is_root ? f.root_norm : subtree_norm(f, lo, hi, level));
    put(p, is_root ? f.n_mults : (size_t)0);
    if (level == 0) {
        for (int q = 0; q < 4; ++q) put(p, (size_t)0);
        if (hi > lo) { memcpy(p, f.tiles.data() + lo * f.tile_bytes, f.tile_bytes); p += f.tile_bytes; }
        return;
    }
    size_t cs[4], cl[4], ce[4];
    size_t q0 = lo;
    for (int q = 0; q < 4; ++q) {
        size_t e = q0;
        while (e < hi && ((f.keys[e] >> (2 * (level - 1))) & 3u) == (uint64_t)q) ++e;
        cl[q] = q0; ce[q] = e;
        cs[q] = e > q0 ? subtree_size(f, q0, e, level - 1) : 0;
        q0 = e;
    }
    for (int q = 0; q < 4; ++q) put(p, cs[q]);
    for (int q = 0; q < 4; ++q)
        if (cs[q]) write_node(f, cl[q], ce[q], level - 1, false, p);
}

Flat flatten(const Matrix& A) {
    Flat f;
    f.b = A.b; f.M = A.M; f.N = A.N; f.depth = A.vdepth();
    f.esize = A.esize(); f.tile_bytes = A.tile_bytes();
    f.root_norm = A.root_norm_cached; f.n_mults = A.n_mults;
    if (A.L) {
        f.keys.resize(A.L); f.norms.resize(A.L * A.esize()); f.tiles.resize(A.L * A.tile_bytes());
        HB_CUDA(cudaMemcpyAsync(f.keys.data(), A.keys.p, A.L * sizeof(uint64_t), cudaMemcpyDeviceToHost, engine().stream));
        HB_CUDA(cudaMemcpyAsync(f.norms.data(), A.norms.p, A.L * A.esize(), cudaMemcpyDeviceToHost, engine().stream));
        HB_CUDA(cudaMemcpyAsync(f.tiles.data(), A.tiles.p, A.L * A.tile_bytes(), cudaMemcpyDeviceToHost, engine().stream));
        sync_stream();
    }
    return f;
}

struct Parsed {
    std::vector<int> bi, bj;
    std::vector<char> norms, tiles;
};

// level = levels remaining below this node (the root carries the matrix's virtual depth, a leaf 0): a buffer whose tree does
// not have the shape the header dimensions imply (H:544-585: nRows halves per level, leaves only at level 0) is rejected
// instead of being read into the wrong block coordinates
void read_node(const char* p, size_t size, size_t esize, int b, int level, uint32_t r, uint32_t c, Parsed& out) {
    const char* end = p + size;
    if (size < header_bytes(esize)) throw_ref("Error in HierarchicalBlockSparseMatrix::assign_from_buffer(): buffer too small.");
    const int nRows = take<int>(p);
    take<int>(p); take<int>(p); take<int>(p);
    const int bs = take<int>(p);
    const char* norm_p = p;
    p += esize;
    take<size_t>(p);
    size_t cs[4];
    for (int q = 0; q < 4; ++q) cs[q] = take<size_t>(p);
    if (bs != b) throw Error(HBSM_E_ARG, "hbsm_b200: assign_from_buffer: inconsistent blocksize inside the buffer");
    if (level < 0 || (long long)nRows != ((long long)b << level))
        throw Error(HBSM_E_ARG, "hbsm_b200: assign_from_buffer: node size does not match its level in the tree");
    bool any = false;
    for (int q = 0; q < 4; ++q) {
        if (!cs[q]) continue;
        any = true;
        if (cs[q] > (size_t)(end - p)) throw_ref("Error in HierarchicalBlockSparseMatrix::assign_from_buffer(): buffer too small.");
        if (level == 0) throw Error(HBSM_E_ARG, "hbsm_b200: assign_from_buffer: a lowest-level node has children");
        read_node(p, cs[q], esize, b, level - 1, 2 * r + (q & 1), 2 * c + ((q >> 1) & 1), out);   // digit = 2*colbit + rowbit
        p += cs[q];
    }
    if (!any && p < end) {   // leaf: the rest is the dense block
        const size_t tb = (size_t)b * b * esize;
        if ((size_t)(end - p) != tb || level != 0)
            throw Error(HBSM_E_ARG, "hbsm_b200: assign_from_buffer: malformed leaf record");
        out.bi.push_back((int)r); out.bj.push_back((int)c);
        out.norms.insert(out.norms.end(), norm_p, norm_p + esize);
        out.tiles.insert(out.tiles.end(), p, p + tb);
    }
}

}  // namespace

size_t serialized_size(const Matrix& A) {
    if (A.empty() || A.L == 0) return header_bytes(A.esize()) + (A.sized && A.vdepth() == 0 ? A.tile_bytes() : 0);
    ensure_engine();
    std::vector<uint64_t> keys = std::vector<uint64_t>(A.L);
    HB_CUDA(cudaMemcpyAsync(keys.data(), A.keys.p, A.L * sizeof(uint64_t), cudaMemcpyDeviceToHost, engine().stream));
    sync_stream();
    Flat f;
    f.b = A.b; f.esize = A.esize(); f.tile_bytes = A.tile_bytes(); f.keys.swap(keys);
    return subtree_size(f, 0, A.L, A.vdepth());
}

void serialize(const Matrix& A, char* buf, size_t cap) {
    if (cap < serialized_size(A)) throw_ref("Error in HierarchicalBlockSparseMatrix<Treal>::write_to_buffer(): buffer too small.");
    char* p = buf;
    if (A.empty()) {   // H:1172-1210: zero dims, then the unset blocksize, norm and counter
        put(p, 0); put(p, 0); put(p, 0); put(p, 0); put(p, A.b);
        put_real(p, A.esize(), 0.0);
        put(p, A.n_mults);
        for (int q = 0; q < 4; ++q) put(p, (size_t)0);
        return;
    }
    ensure_engine();
    Flat f = flatten(A);
    write_node(f, 0, A.L, f.depth, true, p);
}

void deserialize(Matrix& A, const char* buf, size_t size) {
    const size_t es = A.esize();
    if (size < header_bytes(es)) throw_ref("Error in HierarchicalBlockSparseMatrix::assign_from_buffer(): buffer too small.");
    const char* p = buf;
    take<int>(p); take<int>(p);
    const int M = take<int>(p), N = take<int>(p), b = take<int>(p);
    double root_norm;
    if (es == 8) root_norm = take<double>(p); else root_norm = (double)take<float>(p);
    const size_t n_mults = take<size_t>(p);
    A.clear();
    A.b = b;
    A.n_mults = n_mults;
    if (M == 0 && N == 0 && size == header_bytes(es)) return;   // an empty matrix was written
    ensure_engine();
    A.resize(M, N);
    Parsed parsed;
    read_node(buf, size, es, b, A.vdepth(), 0, 0, parsed);
    if (!parsed.bi.empty()) {
        if (A.vdepth() == 0) {
            HB_CUDA(cudaMemcpyAsync(A.tiles.p, parsed.tiles.data(), A.tile_bytes(), cudaMemcpyHostToDevice, engine().stream));
            HB_CUDA(cudaMemcpyAsync(A.norms.p, parsed.norms.data(), es, cudaMemcpyHostToDevice, engine().stream));
            sync_stream();
        } else {
            assign_tiles_host(A, parsed.bi.size(), parsed.bi.data(), parsed.bj.data(), parsed.tiles.data());
            // leaves were written in child order = ascending Morton order = the table's order
            HB_CUDA(cudaMemcpyAsync(A.norms.p, parsed.norms.data(), parsed.bi.size() * es, cudaMemcpyHostToDevice, engine().stream));
            sync_stream();
        }
    }
    A.root_norm_cached = root_norm;
}

}  // namespace hbsm_b200
