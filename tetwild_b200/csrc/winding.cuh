// winding.cuh -- the winding-number hierarchy as it lives in HBM (shared by winding.cu: evaluation + host build, and
// winding_build.cu: device build).
#pragma once
#include "common.cuh"

struct __align__(16) WNode {
    float lo[3], hi[3];      // bounding box, rounded outward
    uint32_t cap_off, cap_cnt;  // cap polyline points [cap_off, cap_off + cap_cnt) of WView::caps
    double apex[3];
    uint32_t tri_off, tri_cnt;  // facets of the whole subtree (contiguous in sorted order)
};
static_assert(sizeof(WNode) == 64, "WNode is 64 bytes");

struct WView {
    const WNode* nodes;    // heap, index 1 .. 2*nBlkP-1
    const double* caps;    // 4 doubles per polyline point: x, y, z, flag (1.0 = first point of a chain)
    const double* tris;    // 9 doubles per facet, sorted
    uint32_t nBlkP;        // leaf blocks, power of two
    uint32_t nF;
};


struct twg_winding {
    twg_ctx* ctx = nullptr;
    WNode* nodes = nullptr;
    double* caps = nullptr;
    double* tris = nullptr;
    uint32_t nBlkP = 1, nF = 0;
    uint64_t n_nodes = 0, n_caps = 0;
    uint32_t leaf = 64;  // triangles per leaf block (TWG_WINDING_LEAF)
    double sort_box[6] = {0, 0, 0, 0, 0, 0};  // surface bbox grown by 10 %: Morton quantisation box of query batches
    bool sort_queries = true;
    std::vector<twg_winding*> replicas;  // handle made on a multi-device context: one replica per device
    WView view() const { return WView{nodes, caps, tris, nBlkP, nF}; }
};


// the hierarchy as three host arrays (host build) / three device arrays (device build)
struct HostTree {
    std::vector<WNode> nodes;
    std::vector<double> caps;
    std::vector<double> tris;
    uint32_t nBlkP = 1;
};
struct DeviceTree {
    WNode* nodes = nullptr;
    double* caps = nullptr;
    double* tris = nullptr;
    uint32_t nBlkP = 1;
    uint64_t n_nodes = 0, n_caps = 0;   // caps: polyline points (4 doubles each)
    double root_lo[3] = {0, 0, 0}, root_hi[3] = {0, 0, 0};
};
// csrc/winding_build.cu: the whole construction on the device, bit-identical to build_host_tree (winding.cu)
int twg_winding_build_device(twg_ctx* c, const double* dV, uint32_t nV, const uint32_t* dF, uint32_t nF, uint32_t kLeaf, DeviceTree* out);
