// Mesh.hh -- the FCC-tetrahedralised brick mesh and its spatial decomposition (host model).
//
// Produces, for the domains owned by one rank, exactly the arrays the reference builds through
// GlobalFccGrid / MeshPartition / MC_Domain (src/GlobalFccGrid.cc, src/MeshPartition.cc,
// src/MC_Domain.cc:77-394) -- same cell order, same node coordinates, same plane coefficients, same
// adjacency -- but directly in the flat per-cell layout the tracking kernels read.  The reference
// discovers a domain's cells with a std::map/std::set flood fill from the domain centre; here every
// global cell is assigned to its nearest centre in one pass (the same nearest-centre rule,
// src/GridAssignmentObject.cc:92-105, ties to the lower domain id), which yields the same partition
// whenever the reference's flood fill succeeds (connected domains).
#ifndef QSB_MESH_HH
#define QSB_MESH_HH

#include <cstdint>
#include <vector>
#include "Parameters.hh"
#include "NuclearData.hh"

namespace qsb {

struct Vec3 { double x, y, z; };

// Global structured grid with 4 node lattices: corners + x-, y-, z-face centres
// (src/GlobalFccGrid.cc:17-29,112-131).
class GlobalFccGrid
{
public:
    GlobalFccGrid(int nx, int ny, int nz, double lx, double ly, double lz);
    int nx, ny, nz;
    double lx, ly, lz, dx, dy, dz;

    int64_t cellGid(int ix, int iy, int iz) const { return ix + (int64_t)nx * (iy + (int64_t)ny * iz); }
    void cellTuple(int64_t gid, int& ix, int& iy, int& iz) const;
    int64_t whichCell(const Vec3& r) const;                 // src/GlobalFccGrid.cc:31-37
    Vec3 cellCenter(int64_t gid) const;                     // src/GlobalFccGrid.cc:39-45
    Vec3 nodeCoord(int ix, int iy, int iz, int basis) const;
    void cellNodes(int64_t gid, Vec3 out[14]) const;        // 8 corners then 6 face centres, src/GlobalFccGrid.cc:48-70
    void faceNeighbors(int64_t gid, int64_t out[6]) const;  // +x -x +y -y +z -z, snapped to self at the boundary
};

// facet f of a cell uses points kFacetPoints[f][0..2] of the cell's 14-point list; the matching facet
// in the face neighbour is kOpposingFacet[f]  (src/MC_Domain.cc:41-50)
extern const int kFacetPoints[24][3];
extern const int kOpposingFacet[24];

// One spatial domain, flat arrays indexed by domain-local cell (ascending global cell id,
// src/MeshPartition.cc:144-155).
struct Domain
{
    int globalDomain = 0;
    int nCells = 0;
    std::vector<int32_t>  cellGid;       // [nCells]
    std::vector<double>   nodes;         // [nCells][14][3]
    std::vector<double>   planes;        // [nCells][24][4]
    std::vector<uint8_t>  faceEvent;     // [nCells][6]
    std::vector<int32_t>  faceAdjCell;   // [nCells][6]  cell index local to the adjacent domain
    std::vector<int32_t>  faceAdjDomain; // [nCells][6]  adjacent domain's index on its rank
    std::vector<int32_t>  faceNbrRank;   // [nCells][6]  owner rank (-1 for boundary / same-rank faces)
    std::vector<int32_t>  faceAdjGlobalDomain; // [nCells][6]
    std::vector<int32_t>  material;      // [nCells]
    std::vector<double>   volume;        // [nCells]
    std::vector<uint64_t> cellId;        // [nCells]  seed base
    std::vector<uint64_t> sourceTally;   // [nCells]  particles sourced so far
};

struct DecompositionInfo
{
    int nRanks = 1, myRank = 0, nDomainsPerRank = 1;
    std::vector<Vec3> centers;           // one per global domain
    std::vector<int> myDomainGids;
    std::vector<double> globalVolumeRate; // nRanks > 1 only: volume * sourceRate of every global cell, by global id
    int rankOf(int domainGid) const { return domainGid / nDomainsPerRank; }      // src/DecompositionObject.cc:41-45
    int indexOf(int domainGid) const { return domainGid % nDomainsPerRank; }
};

// src/initMC.cc:240-320.  Throws std::runtime_error on inconsistent input.
void initMesh(const Parameters& params, const MaterialDatabase& db, int myRank, int nRanks,
              DecompositionInfo& ddc, std::vector<Domain>& domains);

// cell "centre" used for material lookup and coordinate sampling: sum of the 14 points times 1/14
// (src/MCT.cc:231-253)
Vec3 cellPosition(const double* nodes14);

} // namespace qsb
#endif
