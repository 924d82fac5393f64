// orb_vocab.cu — BoW assignment on the GPU: DBoW2 TemplatedVocabulary::transform for ORB descriptors
// (Thirdparty/DBoW2/DBoW2/TemplatedVocabulary.h:1231-1272, FORB::distance FORB.cpp:81-101), the producer of the FeatureVector
// the ORBmatcher searches consume.  One warp per descriptor: the <= 32 children of the current node are scored
// by the lanes in parallel (XOR + POPC on 8 words), a (distance, child order) warp minimum picks the reference's "first child
// with the smallest distance", L levels deep.  The ORB vocabulary (k=10, L=6: 1.1 M nodes x 32 B = 35 MB) is L2-resident.
#include "orb_common.cuh"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <fstream>
#include <sstream>
#include <string>
#include <vector>

struct orbv_vocabulary {
    int k = 0, L = 0, scoring = 0, weighting = 0, nNodes = 0, nWords = 0, device = 0;
    // host copies (kept for save / info)
    std::vector<int> parent;
    std::vector<u8> desc, leaf;
    std::vector<double> weight;
    // device tree: children of node i are childList[childOff[i] .. childOff[i+1])
    int2* d_nodeRec = nullptr;
    bool contiguous = false;
    int maxChildren = 0;
    int *d_childList = nullptr, *d_wordId = nullptr;
    u8* d_desc = nullptr;
    double* d_weight = nullptr;
    cudaStream_t stream = nullptr;
};

// node record: {first child, number of children} when the children of every node are consecutive node ids (true for every
// vocabulary DBoW2 itself created or saved: HKmeansStep appends the k children of a node together), else an index into the
// generic child list.  LPD lanes cooperate on one descriptor (LPD = 16 covers k <= 16 with two descriptors per warp).
template <int LPD, bool CONTIG>
__global__ void __launch_bounds__(256) k_bow_transform(const int2* __restrict__ nodeRec, const int* __restrict__ childList,
                                                       const u8* __restrict__ ndesc, const double* __restrict__ nweight,
                                                       const int* __restrict__ nword, const u8* __restrict__ desc, int n, int L,
                                                       int levelsup, int* __restrict__ word_id, double* __restrict__ weight,
                                                       int* __restrict__ node_id) {
    const int sub = (threadIdx.x & 31) / LPD, sl = threadIdx.x & (LPD - 1);
    const int i = (blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * (32 / LPD) + sub;
    const bool live = i < n;
    const uint4* dp = reinterpret_cast<const uint4*>(desc + (size_t)(live ? i : 0) * 32);
    const uint4 a = __ldg(dp), b = __ldg(dp + 1);
    const int nid_level = L - levelsup;
    int cur = 0, level = 0, nid = 0;
    bool done = !live;
    // The record of the current node travels with the descent: every lane loads its child's record together with the child's
    // descriptor (two independent loads), and the winner's record is shuffled out, so a level costs ONE dependent L2 round trip.
    int2 rec = done ? make_int2(0, 0) : __ldg(nodeRec + 0);
    while (__any_sync(0xffffffffu, !done)) {
        u32 best = 0xFFFFFFFFu;
        const int c0 = rec.x;
        int2 crec = make_int2(0, 0);
        if (!done) {
            const int nc = rec.y;
            if (nc == 0) done = true;                                  // isLeaf()
            else {
                for (int j = sl; j < nc; j += LPD) {
                    const int child = CONTIG ? c0 + j : childList[c0 + j];
                    const uint4* cp = reinterpret_cast<const uint4*>(ndesc + (size_t)child * 32);
                    const uint4 x = __ldg(cp), y = __ldg(cp + 1);
                    const int2 r = __ldg(nodeRec + child);
                    const int d = __popc(a.x ^ x.x) + __popc(a.y ^ x.y) + __popc(a.z ^ x.z) + __popc(a.w ^ x.w) +
                                  __popc(b.x ^ y.x) + __popc(b.y ^ y.y) + __popc(b.z ^ y.z) + __popc(b.w ^ y.w);
                    const u32 key = ((u32)d << 20) | (u32)j;           // strict `d < best_d` in child order == min over (d, order)
                    if (key < best) { best = key; crec = r; }
                }
            }
        }
        const u32 mine = best;
#pragma unroll
        for (int o = LPD / 2; o > 0; o >>= 1) best = min(best, __shfl_xor_sync(0xffffffffu, best, o));
        // the lane that holds the winning key publishes the winner's record (keys are unique inside a group: the order is part of them)
        const u32 grp = (LPD == 32) ? 0xFFFFFFFFu : (((1u << LPD) - 1u) << (sub * LPD));
        const u32 who = __ballot_sync(0xffffffffu, !done && mine == best) & grp;
        const int src = who ? (__ffs(who) - 1) : (int)(threadIdx.x & 31);
        const int rx = __shfl_sync(0xffffffffu, crec.x, src), ry = __shfl_sync(0xffffffffu, crec.y, src);
        if (!done) {
            ++level;
            const int j = (int)(best & 0xFFFFF);
            cur = CONTIG ? c0 + j : childList[c0 + j];
            rec = make_int2(rx, ry);
            if (level == nid_level) nid = cur;
        }
    }
    if (live && sl == 0) {
        if (word_id) word_id[i] = nword[cur];
        if (weight) weight[i] = nweight[cur];
        if (node_id) node_id[i] = nid;
    }
}

static int build_device(orbv_vocabulary* v) {
    const int n = v->nNodes;
    std::vector<int> off(n + 1, 0), list(std::max(n - 1, 1)), word(n, -1);
    for (int i = 1; i < n; i++) {
        ORB_REQUIRE(v->parent[i] >= 0 && v->parent[i] < i, ORB_ERR_ARG, "vocabulary node %d has parent %d (must precede it)", i, v->parent[i]);
        off[v->parent[i] + 1]++;
    }
    for (int i = 0; i < n; i++) off[i + 1] += off[i];
    std::vector<int> fill(off.begin(), off.end() - 1);
    for (int i = 1; i < n; i++) list[fill[v->parent[i]]++] = i;       // children in node (= file) order
    int nw = 0;
    for (int i = 1; i < n; i++)
        if (v->leaf[i]) word[i] = nw++;
    v->nWords = nw;
    // DBoW2 decides "leaf" by children.empty(); the file flag must agree or transform() would index m_words wrongly
    for (int i = 1; i < n; i++)
        ORB_REQUIRE((off[i + 1] == off[i]) == (v->leaf[i] != 0), ORB_ERR_ARG, "vocabulary node %d: leaf flag and child list disagree", i);
    // children consecutive?  then a node record is {first child, count}; else {offset into the child list, count}
    std::vector<int2> rec(n);
    v->contiguous = true;
    v->maxChildren = 0;
    for (int i = 0; i < n; i++) {
        const int c0 = off[i], nc = off[i + 1] - off[i];
        v->maxChildren = std::max(v->maxChildren, nc);
        for (int j = 1; j < nc; j++)
            if (list[c0 + j] != list[c0] + j) v->contiguous = false;
    }
    for (int i = 0; i < n; i++) {
        const int c0 = off[i], nc = off[i + 1] - off[i];
        rec[i] = make_int2(nc == 0 ? 0 : (v->contiguous ? list[c0] : c0), nc);
    }
    ORB_CUDA_TRY(cudaSetDevice(v->device));
    ORB_CUDA_TRY(cudaMalloc(&v->d_nodeRec, n * sizeof(int2)));
    ORB_CUDA_TRY(cudaMalloc(&v->d_childList, list.size() * sizeof(int)));
    ORB_CUDA_TRY(cudaMalloc(&v->d_wordId, n * sizeof(int)));
    ORB_CUDA_TRY(cudaMalloc(&v->d_desc, (size_t)n * 32));
    ORB_CUDA_TRY(cudaMalloc(&v->d_weight, n * sizeof(double)));
    ORB_CUDA_TRY(cudaMemcpy(v->d_nodeRec, rec.data(), n * sizeof(int2), cudaMemcpyHostToDevice));
    ORB_CUDA_TRY(cudaMemcpy(v->d_childList, list.data(), list.size() * sizeof(int), cudaMemcpyHostToDevice));
    ORB_CUDA_TRY(cudaMemcpy(v->d_wordId, word.data(), n * sizeof(int), cudaMemcpyHostToDevice));
    ORB_CUDA_TRY(cudaMemcpy(v->d_desc, v->desc.data(), (size_t)n * 32, cudaMemcpyHostToDevice));
    ORB_CUDA_TRY(cudaMemcpy(v->d_weight, v->weight.data(), n * sizeof(double), cudaMemcpyHostToDevice));
    ORB_CUDA_TRY(cudaStreamCreateWithFlags(&v->stream, cudaStreamNonBlocking));
    return ORB_OK;
}

extern "C" void orbv_destroy(orbv_vocabulary* v) {
    if (!v) return;
    cudaSetDevice(v->device);
    cudaFree(v->d_nodeRec); cudaFree(v->d_childList); cudaFree(v->d_wordId); cudaFree(v->d_desc); cudaFree(v->d_weight);
    if (v->stream) cudaStreamDestroy(v->stream);
    delete v;
}

static int finish_create(orbv_vocabulary* v, orbv_vocabulary** out) {
    ORB_REQUIRE(orb_device_count() > v->device && v->device >= 0, ORB_ERR_CUDA, "CUDA device %d not available (no CPU fallback)", v->device);
    const int rc = build_device(v);
    if (rc != ORB_OK) { orbv_destroy(v); return rc; }
    *out = v;
    return ORB_OK;
}

extern "C" int orbv_create(orbv_vocabulary** out, int k, int L, int scoring, int weighting, int n_nodes, const int* parent,
                           const uint8_t* descriptors, const double* weights, const uint8_t* is_leaf, int device) {
    ORB_REQUIRE(out && parent && descriptors && weights && is_leaf && n_nodes >= 1 && k >= 1 && L >= 1, ORB_ERR_ARG, "bad arguments");
    *out = nullptr;
    orbv_vocabulary* v = new orbv_vocabulary();
    v->k = k; v->L = L; v->scoring = scoring; v->weighting = weighting; v->nNodes = n_nodes; v->device = device;
    v->parent.assign(parent, parent + n_nodes);
    v->desc.assign(descriptors, descriptors + (size_t)n_nodes * 32);
    v->weight.assign(weights, weights + n_nodes);
    v->leaf.assign(is_leaf, is_leaf + n_nodes);
    v->leaf[0] = n_nodes == 1;
    return finish_create(v, out);
}

// TemplatedVocabulary::loadFromTextFile (:1351-1440): "k L scoring weighting" then one node per line:
// "parent isLeaf d0 ... d31 weight".  (The reference also turns a trailing empty line into a garbage node; that is not reproduced.)
extern "C" int orbv_load_text(orbv_vocabulary** out, const char* path, int device) {
    ORB_REQUIRE(out && path, ORB_ERR_ARG, "bad arguments");
    *out = nullptr;
    std::ifstream f(path);
    ORB_REQUIRE(f.good(), ORB_ERR_ARG, "cannot open %s", path);
    std::string line;
    std::getline(f, line);
    int k = -1, L = -1, n1 = -1, n2 = -1;
    { std::stringstream ss(line); ss >> k >> L >> n1 >> n2; }
    ORB_REQUIRE(!(k < 0 || k > 20 || L < 1 || L > 10 || n1 < 0 || n1 > 5 || n2 < 0 || n2 > 3), ORB_ERR_ARG,
                "Vocabulary loading failure: This is not a correct text file!");
    orbv_vocabulary* v = new orbv_vocabulary();
    v->k = k; v->L = L; v->scoring = n1; v->weighting = n2; v->device = device;
    v->parent.push_back(0); v->leaf.push_back(0); v->weight.push_back(0.0); v->desc.resize(32, 0);
    while (std::getline(f, line)) {
        if (line.find_first_not_of(" \t\r\n") == std::string::npos) continue;
        std::stringstream ss(line);
        int pid = 0, isLeaf = 0;
        ss >> pid >> isLeaf;
        u8 d[32];
        for (int i = 0; i < 32; i++) { int x = 0; ss >> x; d[i] = (u8)x; }
        double w = 0;
        ss >> w;
        v->parent.push_back(pid); v->leaf.push_back(isLeaf > 0 ? 1 : 0); v->weight.push_back(w);
        v->desc.insert(v->desc.end(), d, d + 32);
    }
    v->nNodes = (int)v->parent.size();
    return finish_create(v, out);
}

// TemplatedVocabulary::loadFromBinaryFile (:1467-1512): u32 nb_nodes, u32 size_node, int k, int L, int scoring, int weighting,
// then nb_nodes-1 records {int parent; u8 desc[32]; float weight; bool is_leaf}.
extern "C" int orbv_load_binary(orbv_vocabulary** out, const char* path, int device) {
    ORB_REQUIRE(out && path, ORB_ERR_ARG, "bad arguments");
    *out = nullptr;
    std::ifstream f(path, std::ios::binary);
    ORB_REQUIRE(f.good(), ORB_ERR_ARG, "cannot open %s", path);
    unsigned nb = 0, sz = 0;
    int k = 0, L = 0, sc = 0, wt = 0;
    f.read((char*)&nb, 4); f.read((char*)&sz, 4); f.read((char*)&k, 4); f.read((char*)&L, 4); f.read((char*)&sc, 4); f.read((char*)&wt, 4);
    ORB_REQUIRE(f.good() && sz == 4 + 32 + 4 + 1 && nb >= 1, ORB_ERR_ARG, "%s is not a binary ORB vocabulary (size_node %u)", path, sz);
    orbv_vocabulary* v = new orbv_vocabulary();
    v->k = k; v->L = L; v->scoring = sc; v->weighting = wt; v->device = device; v->nNodes = (int)nb;
    v->parent.assign(nb, 0); v->leaf.assign(nb, 0); v->weight.assign(nb, 0.0); v->desc.assign((size_t)nb * 32, 0);
    std::vector<char> buf(sz);
    for (unsigned i = 1; i < nb; i++) {
        f.read(buf.data(), sz);
        if (!f.good()) { delete v; orb_set_error("%s: truncated at node %u", path, i); return ORB_ERR_ARG; }
        int pid; float w;
        memcpy(&pid, buf.data(), 4);
        memcpy(&v->desc[(size_t)i * 32], buf.data() + 4, 32);
        memcpy(&w, buf.data() + 36, 4);
        v->parent[i] = pid; v->weight[i] = (double)w; v->leaf[i] = buf[40] ? 1 : 0;
    }
    return finish_create(v, out);
}

extern "C" int orbv_save_binary(const orbv_vocabulary* v, const char* path) {
    ORB_REQUIRE(v && path, ORB_ERR_ARG, "bad arguments");
    std::ofstream f(path, std::ios::binary);
    ORB_REQUIRE(f.good(), ORB_ERR_ARG, "cannot open %s", path);
    const unsigned nb = (unsigned)v->nNodes, sz = 4 + 32 + 4 + 1;
    f.write((const char*)&nb, 4); f.write((const char*)&sz, 4); f.write((const char*)&v->k, 4); f.write((const char*)&v->L, 4);
    f.write((const char*)&v->scoring, 4); f.write((const char*)&v->weighting, 4);
    for (unsigned i = 1; i < nb; i++) {
        const float w = (float)v->weight[i];
        const char leaf = v->leaf[i] ? 1 : 0;
        f.write((const char*)&v->parent[i], 4); f.write((const char*)&v->desc[(size_t)i * 32], 32); f.write((const char*)&w, 4); f.write(&leaf, 1);
    }
    ORB_REQUIRE(f.good(), ORB_ERR_ARG, "write to %s failed", path);
    return ORB_OK;
}

extern "C" int orbv_info(const orbv_vocabulary* v, int* k, int* L, int* n_nodes, int* n_words, int* scoring, int* weighting) {
    ORB_REQUIRE(v, ORB_ERR_ARG, "null vocabulary");
    if (k) *k = v->k;
    if (L) *L = v->L;
    if (n_nodes) *n_nodes = v->nNodes;
    if (n_words) *n_words = v->nWords;
    if (scoring) *scoring = v->scoring;
    if (weighting) *weighting = v->weighting;
    return ORB_OK;
}

extern "C" int orbv_transform_device(const orbv_vocabulary* v, const uint8_t* d_desc, int n, int levelsup, int* d_word_id,
                                     double* d_weight, int* d_node_id, void* stream) {
    ORB_REQUIRE(v && (d_desc || n == 0) && n >= 0, ORB_ERR_ARG, "bad arguments");
    if (n == 0) return ORB_OK;
    cudaStream_t st = stream ? (cudaStream_t)stream : v->stream;
#define BOW_LAUNCH(LPD, CT)                                                                                                   \
    k_bow_transform<LPD, CT><<<orb_div_up(n, 8 * (32 / LPD)), 256, 0, st>>>(v->d_nodeRec, v->d_childList, v->d_desc, v->d_weight, \
                                                                             v->d_wordId, d_desc, n, v->L, levelsup, d_word_id,    \
                                                                             d_weight, d_node_id)
    if (v->maxChildren <= 16) { if (v->contiguous) BOW_LAUNCH(16, true); else BOW_LAUNCH(16, false); }
    else { if (v->contiguous) BOW_LAUNCH(32, true); else BOW_LAUNCH(32, false); }
#undef BOW_LAUNCH
    ORB_CUDA_TRY(cudaGetLastError());
    return ORB_OK;
}

extern "C" int orbv_transform(const orbv_vocabulary* v, const uint8_t* desc, int n, int levelsup, int* word_id, double* weight, int* node_id) {
    ORB_REQUIRE(v && (desc || n == 0) && n >= 0, ORB_ERR_ARG, "bad arguments");
    if (n == 0) return ORB_OK;
    ORB_CUDA_TRY(cudaSetDevice(v->device));
    u8* d_desc = nullptr; int *d_w = nullptr, *d_n = nullptr; double* d_wt = nullptr;
    ORB_CUDA_TRY(cudaMallocAsync(&d_desc, (size_t)n * 32, v->stream));
    ORB_CUDA_TRY(cudaMallocAsync(&d_w, (size_t)n * 4, v->stream));
    ORB_CUDA_TRY(cudaMallocAsync(&d_n, (size_t)n * 4, v->stream));
    ORB_CUDA_TRY(cudaMallocAsync(&d_wt, (size_t)n * 8, v->stream));
    ORB_CUDA_TRY(cudaMemcpyAsync(d_desc, desc, (size_t)n * 32, cudaMemcpyHostToDevice, v->stream));
    int rc = orbv_transform_device(v, d_desc, n, levelsup, d_w, d_wt, d_n, v->stream);
    if (rc == ORB_OK) {
        if (word_id) ORB_CUDA_TRY(cudaMemcpyAsync(word_id, d_w, (size_t)n * 4, cudaMemcpyDeviceToHost, v->stream));
        if (weight) ORB_CUDA_TRY(cudaMemcpyAsync(weight, d_wt, (size_t)n * 8, cudaMemcpyDeviceToHost, v->stream));
        if (node_id) ORB_CUDA_TRY(cudaMemcpyAsync(node_id, d_n, (size_t)n * 4, cudaMemcpyDeviceToHost, v->stream));
    }
    cudaFreeAsync(d_desc, v->stream); cudaFreeAsync(d_w, v->stream); cudaFreeAsync(d_n, v->stream); cudaFreeAsync(d_wt, v->stream);
    ORB_CUDA_TRY(cudaStreamSynchronize(v->stream));
    return rc;
}
