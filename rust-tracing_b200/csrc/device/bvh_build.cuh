// BVHNode::node_from_list (bvh.rs:31-66) on the device, with the reference's topology rule: node i draws an axis, the
// objects of its range are sorted by the minimum of their boxes on that axis (ties keep their order, as the host build's
// stable sort with the crate's strict `<` comparator does), the range splits at span / 2; two objects form a pair that
// swaps unless the first is strictly less; one object is a leaf.
//
// Splitting at span / 2 makes the SHAPE of the tree (who covers which range, pre-order numbering, how many axis draws) a
// function of n alone, and the axis draws a function of the seed: both are laid out on the host in microseconds. What
// costs time on the host is the sorting - n log n comparisons per level over the crate's trait objects - and that is
// what runs here, level by level, as a rank sort: every element counts the elements of its range that come before it
// (keys staged through shared memory in tiles) and writes itself to that position. O(span^2) per range, exact with
// respect to the comparator, no data-dependent control flow; ~2 n^2 comparisons in total (n = 1000: 2 M).
// Node boxes are unions of leaf boxes (fmin / fmax: exact and order-independent).
//
// Included by rt_cuda.cu.
#pragma once

namespace {

struct BvhCall {            // one call of node_from_list
    int start, span;        // range of the (current) object order it covers
    int level;
    int node;               // pre-order index of the node it creates
    int axis;
};

// keys: box minimum on the range's axis. seg_* are per ELEMENT for this level (span 0 = not sorted at this level).
__global__ void bvh_rank_sort_kernel(const double* __restrict__ bbox, const int* __restrict__ perm_in, int* __restrict__ perm_out,
                                     const int* __restrict__ seg_start, const int* __restrict__ seg_span,
                                     const int* __restrict__ seg_axis, int n) {
    __shared__ double tile[256];
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    // a block may straddle ranges: every thread walks ITS range; tiles are loaded cooperatively per distinct range in turn
    int my_start = 0, my_span = 0, my_axis = 0, my_id = -1;
    double my_key = 0.0;
    if (p < n) {
        my_start = seg_start[p]; my_span = seg_span[p]; my_axis = seg_axis[p];
        my_id = perm_in[p];
        if (my_span > 0) my_key = bbox[(size_t)my_id * 6 + 2 * my_axis];
    }
    // ranges touched by this block: from the first thread's range to the last thread's range (ranges are contiguous)
    const int first_p = blockIdx.x * blockDim.x;
    const int last_p = min(n, first_p + (int)blockDim.x) - 1;
    int rank = 0;
    int r_start = seg_start[first_p];
    const int stop = seg_start[last_p] + max(seg_span[last_p], 1);
    while (r_start < stop) {
        const int r_span = max(seg_span[r_start], 1), r_axis = seg_axis[r_start];
        const bool sorted_range = seg_span[r_start] > 0;
        if (sorted_range) {
            for (int base = r_start; base < r_start + r_span; base += 256) {
                const int q = base + threadIdx.x;
                __syncthreads();
                if (q < r_start + r_span) tile[threadIdx.x] = bbox[(size_t)perm_in[q] * 6 + 2 * r_axis];
                __syncthreads();
                if (my_span > 0 && my_start == r_start) {
                    const int m = min(256, r_start + r_span - base);
                    for (int k = 0; k < m; ++k) {
                        const double key = tile[k];
                        rank += (key < my_key) || (key == my_key && base + k < p);
                    }
                }
            }
        }
        r_start += r_span;
    }
    if (p < n) perm_out[my_span > 0 ? my_start + rank : p] = my_id;
}

// the pairs (bvh.rs:45-57): swap unless strictly less
__global__ void bvh_pair_kernel(const double* __restrict__ bbox, int* __restrict__ perm, const BvhCall* __restrict__ calls, int n_calls) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n_calls || calls[c].span != 2) return;
    const int a = perm[calls[c].start], b = perm[calls[c].start + 1];
    const int ax = calls[c].axis;
    if (!(bbox[(size_t)a * 6 + 2 * ax] < bbox[(size_t)b * 6 + 2 * ax])) { perm[calls[c].start] = b; perm[calls[c].start + 1] = a; }
}

// one WARP per call: the box of its range (lanes stride over the range, fmin / fmax butterflies; exact and order-independent)
__global__ void bvh_box_kernel(const double* __restrict__ bbox, const int* __restrict__ perm, const BvhCall* __restrict__ calls,
                               int n_calls, double* __restrict__ out) {
    const int c = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (c >= n_calls) return;
    const int start = calls[c].start, span = calls[c].span;
    const double inf = __longlong_as_double(0x7ff0000000000000ll);
    double b[6] = {inf, -inf, inf, -inf, inf, -inf};
    for (int j = lane; j < span; j += 32) {
        const double* g = bbox + (size_t)perm[start + j] * 6;
        for (int k = 0; k < 3; ++k) { b[2 * k] = fmin(b[2 * k], g[2 * k]); b[2 * k + 1] = fmax(b[2 * k + 1], g[2 * k + 1]); }
    }
    for (int off = 16; off > 0; off >>= 1)
        for (int k = 0; k < 3; ++k) {
            b[2 * k] = fmin(b[2 * k], __shfl_xor_sync(0xffffffffu, b[2 * k], off));
            b[2 * k + 1] = fmax(b[2 * k + 1], __shfl_xor_sync(0xffffffffu, b[2 * k + 1], off));
        }
    double v = b[0];
#pragma unroll
    for (int k = 1; k < 6; ++k) if (lane == k) v = b[k];
    if (lane < 6) out[(size_t)c * 6 + lane] = v;
}

void bvh_shape(int start, int span, int level, const int32_t* axes, int* next_axis, int* next_node, std::vector<BvhCall>* calls) {
    BvhCall c{start, span, level, (*next_node)++, axes[(*next_axis)++]};
    calls->push_back(c);
    if (span == 2) { *next_node += 2; return; }
    if (span <= 1) return;
    bvh_shape(start, span / 2, level + 1, axes, next_axis, next_node, calls);
    bvh_shape(start + span / 2, span - span / 2, level + 1, axes, next_axis, next_node, calls);
}

}  // namespace

extern "C" int rt_bvh_axis_draws(int n) {          // how many axis draws node_from_list makes for n objects
    if (n <= 0) return 0;
    if (n <= 2) return 1;
    return 1 + rt_bvh_axis_draws(n / 2) + rt_bvh_axis_draws(n - n / 2);
}

extern "C" int rt_bvh_build_device(rt_context* c, const double* bboxes, int n, const int32_t* axes, rt_bvh_node_desc* nodes_out,
                                   int32_t* order_out) {
    if (!c || !bboxes || !axes || !nodes_out || n <= 0) return fail(RT_ERR_INVALID_ARGUMENT, "rt_bvh_build_device: null argument");
    CU(cudaSetDevice(c->device));
    std::vector<BvhCall> calls;
    int next_axis = 0, next_node = 0;
    bvh_shape(0, n, 0, axes, &next_axis, &next_node, &calls);
    int levels = 0;
    for (const BvhCall& k : calls) levels = std::max(levels, k.level + 1);
    double* d_bbox = nullptr; double* d_boxes = nullptr;
    int *d_perm[2] = {nullptr, nullptr}, *d_seg = nullptr;
    BvhCall* d_calls = nullptr;
    struct Free { std::vector<void*> v; ~Free() { for (void* p : v) cudaFree(p); } } guard;
    auto alloc = [&](void** p, size_t bytes) { cudaError_t e = cudaMalloc(p, bytes ? bytes : 16); if (e == cudaSuccess) guard.v.push_back(*p); return e; };
    CU(alloc((void**)&d_bbox, (size_t)n * 48));
    CU(alloc((void**)&d_perm[0], (size_t)n * 4));
    CU(alloc((void**)&d_perm[1], (size_t)n * 4));
    CU(alloc((void**)&d_seg, (size_t)n * 12));
    CU(alloc((void**)&d_calls, calls.size() * sizeof(BvhCall)));
    CU(alloc((void**)&d_boxes, calls.size() * 48));
    CU(cudaMemcpy(d_bbox, bboxes, (size_t)n * 48, cudaMemcpyHostToDevice));
    CU(cudaMemcpy(d_calls, calls.data(), calls.size() * sizeof(BvhCall), cudaMemcpyHostToDevice));
    std::vector<int> ident((size_t)n);
    for (int i = 0; i < n; ++i) ident[i] = i;
    CU(cudaMemcpy(d_perm[0], ident.data(), (size_t)n * 4, cudaMemcpyHostToDevice));
    int cur = 0;
    std::vector<int> seg((size_t)n * 3);
    const unsigned blocks = (unsigned)((n + 255) / 256);
    for (int L = 0; L < levels; ++L) {
        // per element: the range that is sorted at this level (span 0: none - the element's call ended above, or is a pair / leaf)
        for (int i = 0; i < n; ++i) { seg[i] = i; seg[n + i] = 0; seg[2 * n + i] = 0; }
        bool any = false;
        for (const BvhCall& k : calls) {
            if (k.level != L || k.span <= 2) continue;
            any = true;
            for (int i = 0; i < k.span; ++i) { seg[k.start + i] = k.start; seg[n + k.start + i] = k.span; seg[2 * n + k.start + i] = k.axis; }
        }
        if (!any) continue;
        CU(cudaMemcpy(d_seg, seg.data(), (size_t)n * 12, cudaMemcpyHostToDevice));
        bvh_rank_sort_kernel<<<blocks, 256>>>(d_bbox, d_perm[cur], d_perm[cur ^ 1], d_seg, d_seg + n, d_seg + 2 * n, n);
        CU(cudaGetLastError());
        cur ^= 1;
    }
    const unsigned cblocks = (unsigned)((calls.size() + 127) / 128);
    bvh_pair_kernel<<<cblocks, 128>>>(d_bbox, d_perm[cur], d_calls, (int)calls.size());
    bvh_box_kernel<<<(unsigned)((calls.size() * 32 + 127) / 128), 128>>>(d_bbox, d_perm[cur], d_calls, (int)calls.size(), d_boxes);
    CU(cudaGetLastError());
    std::vector<int> order((size_t)n);
    std::vector<double> boxes(calls.size() * 6);
    CU(cudaMemcpy(order.data(), d_perm[cur], (size_t)n * 4, cudaMemcpyDeviceToHost));
    CU(cudaMemcpy(boxes.data(), d_boxes, boxes.size() * 8, cudaMemcpyDeviceToHost));
    // pre-order node array: topology from the shape, leaf objects from the device's order, boxes from the device
    for (size_t ci = 0; ci < calls.size(); ++ci) {
        const BvhCall& k = calls[ci];
        rt_bvh_node_desc& nd = nodes_out[k.node];
        std::memcpy(nd.bbox, &boxes[ci * 6], 48);
        nd.axis = k.axis;
        if (k.span == 1) {
            nd.left = nd.right = -1;
            nd.object = order[k.start];
        } else if (k.span == 2) {
            nd.left = k.node + 1; nd.right = k.node + 2; nd.object = -1;
            for (int j = 0; j < 2; ++j) {
                rt_bvh_node_desc& lf = nodes_out[k.node + 1 + j];
                std::memcpy(lf.bbox, bboxes + (size_t)order[k.start + j] * 6, 48);
                lf.left = lf.right = -1; lf.object = order[k.start + j]; lf.axis = -1;
            }
        } else {
            nd.object = -1;
            nd.left = k.node + 1;                       // pre-order: the left child follows its parent,
            nd.right = k.node + 2 * (k.span / 2);       // the right one follows the 2 * (span / 2) - 1 nodes of the left subtree
        }
    }
    if (order_out) for (int i = 0; i < n; ++i) order_out[i] = order[i];
    return next_node;
}
