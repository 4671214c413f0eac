// Geometry of one extractor handle: everything the kernels need to know about the pyramid, the FAST cell grid
// and the per-level output slots.  Built once on the host at eaof_orb_create (the reference recomputes these per
// frame: src/ORBextractor.cc:1107-1118 level sizes, :773-787 cell grid, :543-545 root nodes) and passed to the
// kernels by value as a __grid_constant__ parameter.
#pragma once
#include <stdint.h>

#define EAOF_MAX_LEVELS 16
#define EAOF_EDGE 19       // EDGE_THRESHOLD, src/ORBextractor.cc:74
#define EAOF_INNER_X0 32   // byte column of inner pixel x=0 inside a padded level row (16 B aligned)
#define EAOF_MIN_BORDER 16 // EDGE_THRESHOLD-3, src/ORBextractor.cc:773

struct LevelGeom {
    int w, h;            // inner size (mvImagePyramid[l].cols/rows)
    int pitch;           // bytes per padded row (multiple of 64)
    int rows;            // h + 38
    uint32_t off;        // byte offset of the bordered buffer's first row inside one frame's pyramid block
    int nCols, nRows;    // FAST cell grid (0 when the detection window is smaller than one cell)
    int wCell, hCell;
    int winW, winH;      // maxBorderX-minBorderX, maxBorderY-minBorderY
    int nIni;            // root nodes, src/ORBextractor.cc:543
    float hX;            // src/ORBextractor.cc:545
    int quota;           // mnFeaturesPerLevel[l]
    int nodeCap;         // upper bound of the quadtree list size = max(quota+3, 4*nIni)
    uint32_t candOff;    // first candidate slot of this level inside one frame's candidate block
    uint32_t candCap;
    int slotOff;         // first keypoint slot of this level inside one frame's slot block
    int cellOff;         // first entry of this level in the cell table
    int xTab, yTab;      // offsets into the resize coefficient table (ints): [xofs | alpha] per column, [yofs | beta] per row
    float scale;         // mvScaleFactor[l]
    float kpSize;        // (float)(int)(31*scale)
    int blurTaskOff;     // first blur task (column word x row chunk) of this level inside one frame
};

struct Geom {
    int nlevels;
    int W, H;
    int iniTh, minTh;
    int blurMode;
    int cellsPerFrame;
    int slotsPerFrame;    // sum of nodeCap
    int maxNodeCap;
    int blurTasksPerFrame;
    int fastPW;           // k_fast: tile pitch in words (1 pad word + widest cell row), odd
    int fastMapWords;     // k_fast: words of one tile / score map (multiple of 4)
    int fastWarpWords;    // k_fast: shared-memory words per warp (tile + score map + pair list)
    uint32_t candPerFrame;
    uint64_t pyrFrameBytes;  // multiple of 256
    LevelGeom L[EAOF_MAX_LEVELS];
};

struct alignas(16) CellDesc {  // one FAST cell, src/ORBextractor.cc:789-806; 16 bytes: one vector load
    short level;
    short iniX, iniY;  // top-left of the cell sub-image, inner level coordinates
    short cw, ch;      // sub-image size (<= 66)
    short pad, pad2, pad3;
};
