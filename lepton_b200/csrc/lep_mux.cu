// lep_mux.cu -- sm_100a gather kernel that assembles the final .lep files on the device.
//
// Reference: write_ujpg (src/lepton/jpgcoder.cc:3779-4097) writes the fixed header, the compressed JPEG header and
// "CMP"; vp8_full_encoder's tail (src/lepton/vp8_encoder.cc:573-614) interleaves the thread-segment streams through
// MuxWriter (src/io/MuxReader.hh:336-522) and appends the LE32 file size.  The MuxWriter's decisions depend only on the
// stream LENGTHS, so the host runs it data-free (plan_mux, lep_container.cc) once the range coder's lengths are back and
// hands the device a list of pieces: "these header bytes, then `len` bytes from there, at that offset of the output".
// HBM-bound byte moving: one warp per piece, 16-byte stores to aligned destinations, sources at any alignment (a mux
// packet lands wherever the previous one ended) read as aligned words and funnel-shifted into place.
#include <vector>

#include "../../include/lepton_b200.h"
#include "lep_common.cuh"

namespace lepb200 {

struct GatherPiece {
    unsigned long long src;          // device address of the payload bytes (any alignment); readable down to src & ~3 and up to the
                                     // word that holds the last byte (the arenas are 256-byte aligned and padded)
    unsigned long long dst;          // offset in the output buffer of the piece's first header byte
    uint32_t len;                    // payload bytes
    uint8_t nhdr, hdr[3];            // literal bytes in front of the payload (mux packet header: 1 or 3 bytes)
};

// Host side: the pieces of ONE file.  `lit` = address (as the kernel sees it) of the file's literal bytes: its header
// (hdr_len bytes: fixed header, compressed JPEG header, "CMP") followed by room for the 4 trailer bytes, which are
// written here through `lit_host` (the same bytes as the host sees them).  stream_addr[id] = address of stream id.
// Returns the file size; the file starts at offset file_off of the output buffer.
inline uint32_t gather_file_pieces(const lepb200_mux_packet* plan, size_t nplan, const unsigned long long* stream_addr, unsigned long long lit,
                                   uint8_t* lit_host, size_t hdr_len, size_t file_off, std::vector<GatherPiece>& pieces) {
    GatherPiece g;
    memset(&g, 0, sizeof(g));
    g.src = lit; g.dst = file_off; g.len = (uint32_t)hdr_len;
    pieces.push_back(g);
    size_t pos = file_off + hdr_len;
    for (size_t k = 0; k < nplan; ++k) {
        const lepb200_mux_packet& p = plan[k];
        g.src = stream_addr[p.id] + p.src_off; g.dst = pos; g.len = p.len;
        g.nhdr = p.nhdr; g.hdr[0] = p.hdr[0]; g.hdr[1] = p.hdr[1]; g.hdr[2] = p.hdr[2];
        pieces.push_back(g);
        pos += (size_t)p.nhdr + p.len;
    }
    const uint32_t fsz = (uint32_t)(pos + 4 - file_off);                 // LE32 total file size (vp8_encoder.cc:603-614)
    uint8_t* tr = lit_host + hdr_len;
    tr[0] = (uint8_t)fsz; tr[1] = (uint8_t)(fsz >> 8); tr[2] = (uint8_t)(fsz >> 16); tr[3] = (uint8_t)(fsz >> 24);
    memset(&g, 0, sizeof(g));
    g.src = lit + hdr_len; g.dst = pos; g.len = 4;
    pieces.push_back(g);
    return fsz;
}

constexpr int GATHER_WARPS = 8;

__global__ void __launch_bounds__(GATHER_WARPS * 32)
lep_gather_kernel(const GatherPiece* __restrict__ pieces, uint32_t npieces, uint8_t* __restrict__ out) {
    const int lane = threadIdx.x & 31;
    const uint32_t warp0 = blockIdx.x * GATHER_WARPS + (threadIdx.x >> 5), nwarps = gridDim.x * GATHER_WARPS;
    for (uint32_t pi = warp0; pi < npieces; pi += nwarps) {
        const GatherPiece pc = pieces[pi];
        uint8_t* dst = out + pc.dst;
        if (lane < pc.nhdr) dst[lane] = pc.hdr[lane];
        dst += pc.nhdr;
        const uint8_t* src = reinterpret_cast<const uint8_t*>(pc.src);
        uint32_t len = pc.len;
        // head: bytes up to the first 16-byte boundary of the destination
        const uint32_t head = min(len, (uint32_t)((16u - (uint32_t)(reinterpret_cast<uintptr_t>(dst) & 15u)) & 15u));
        if ((uint32_t)lane < head) dst[lane] = src[lane];
        dst += head; src += head; len -= head;
        // body: 16 bytes per lane and step; the source is read as aligned 32-bit words and shifted into place
        const uint32_t n16 = len / 16;
        const uint32_t sh = (uint32_t)(reinterpret_cast<uintptr_t>(src) & 3u) * 8u;
        const uint32_t* sw = reinterpret_cast<const uint32_t*>(reinterpret_cast<uintptr_t>(src) & ~uintptr_t(3));
        uint4* d4 = reinterpret_cast<uint4*>(dst);
        for (uint32_t i = lane; i < n16; i += 32) {
            const uint32_t* w = sw + 4 * i;
            const uint32_t w0 = __ldg(w), w1 = __ldg(w + 1), w2 = __ldg(w + 2), w3 = __ldg(w + 3);
            uint4 v;
            if (sh == 0) { v.x = w0; v.y = w1; v.z = w2; v.w = w3; }
            else {
                const uint32_t w4 = __ldg(w + 4);
                v.x = __funnelshift_r(w0, w1, sh); v.y = __funnelshift_r(w1, w2, sh);
                v.z = __funnelshift_r(w2, w3, sh); v.w = __funnelshift_r(w3, w4, sh);
            }
            d4[i] = v;
        }
        // tail: fewer than 16 bytes
        const uint32_t done = n16 * 16;
        if (done + (uint32_t)lane < len) dst[done + lane] = src[done + lane];
    }
}

}  // namespace lepb200
