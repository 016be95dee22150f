// STL reader — replaces STLReader::readFile / ReadBinary (core/read_STL.hpp:26-77, 131-174).
// Host-only: the file is I/O bound and a few hundred KB; triangles go to the GPU as they are.
#include "wr_common.cuh"

extern "C" int wr_stl_parse(const uint8_t* buf, size_t len, float* tris12, int cap, int* ntri)
{
    WR_REQUIRE(buf && ntri, WR_ERR_INVALID, "wr_stl_parse: null");
    *ntri = 0;
    WR_REQUIRE(len >= 84, WR_ERR_FORMAT, "wr_stl_parse: shorter than a binary STL header");
    // read_STL.hpp:65 — the reference treats the file as binary iff byte 79 of the header is NUL;
    // its ASCII branch (:99-129) never parses normals and is not supported here.
    WR_REQUIRE(buf[79] == '\0', WR_ERR_FORMAT, "wr_stl_parse: ASCII STL is not supported");
    uint32_t n;
    memcpy(&n, buf + 80, 4);   // cpyint :158-165
    WR_REQUIRE(n < 0x7fffffffu && 84 + (size_t)n * 50 <= len, WR_ERR_FORMAT, "wr_stl_parse: triangle count exceeds file size");
    *ntri = (int)n;
    if (!tris12) return WR_OK;
    const uint8_t* p = buf + 84;
    for (uint32_t i = 0; i < n && (int)i < cap; i++, p += 50) memcpy(tris12 + 12 * (size_t)i, p, 48);   // normal + 3 vertices; 2 attribute bytes skipped (:151)
    return WR_OK;
}
