// Mesh.cc -- see Mesh.hh.  Floating-point expressions keep the reference's operation order so every
// coordinate, plane coefficient and volume carries the same bits (checked against oracle/ref_dump.cc).
#include "Mesh.hh"

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <set>
#include <stdexcept>
#include <tuple>

namespace qsb {

const int kFacetPoints[24][3] = {
    {1, 3, 8},  {3, 7, 8},  {7, 5, 8},  {5, 1, 8},       // +x face (centre = point 8)
    {0, 4, 9},  {4, 6, 9},  {6, 2, 9},  {2, 0, 9},       // -x
    {3, 2, 10}, {2, 6, 10}, {6, 7, 10}, {7, 3, 10},      // +y
    {0, 1, 11}, {1, 5, 11}, {5, 4, 11}, {4, 0, 11},      // -y
    {4, 5, 12}, {5, 7, 12}, {7, 6, 12}, {6, 4, 12},      // +z
    {0, 2, 13}, {2, 3, 13}, {3, 1, 13}, {1, 0, 13} };    // -z
const int kOpposingFacet[24] = { 7, 6, 5, 4, 3, 2, 1, 0, 12, 15, 14, 13, 8, 11, 10, 9, 20, 23, 22, 21, 16, 19, 18, 17 };

GlobalFccGrid::GlobalFccGrid(int nx_, int ny_, int nz_, double lx_, double ly_, double lz_)
: nx(nx_), ny(ny_), nz(nz_), lx(lx_), ly(ly_), lz(lz_)
{
    if (nx < 1 || ny < 1 || nz < 1) throw std::runtime_error("mesh needs nx, ny, nz >= 1");
    dx = lx / nx; dy = ly / ny; dz = lz / nz;
}

void GlobalFccGrid::cellTuple(int64_t gid, int& ix, int& iy, int& iz) const
{
    ix = (int)(gid % nx); gid /= nx;
    iy = (int)(gid % ny);
    iz = (int)(gid / ny);
}

int64_t GlobalFccGrid::whichCell(const Vec3& r) const
{
    const int ix = (int)(r.x / dx), iy = (int)(r.y / dy), iz = (int)(r.z / dz);
    return cellGid(ix, iy, iz);
}

Vec3 GlobalFccGrid::nodeCoord(int ix, int iy, int iz, int basis) const
{
    // lattice point + basis offset; basis 0 = corner, 1/2/3 = centre of the x-/y-/z-normal face
    static const int half[4][3] = { {0, 0, 0}, {0, 1, 1}, {1, 0, 1}, {1, 1, 0} };
    Vec3 r = { ix * dx, iy * dy, iz * dz };
    r.x = r.x + (half[basis][0] ? dx / 2.0 : 0.);
    r.y = r.y + (half[basis][1] ? dy / 2.0 : 0.);
    r.z = r.z + (half[basis][2] ? dz / 2.0 : 0.);
    return r;
}

Vec3 GlobalFccGrid::cellCenter(int64_t gid) const
{
    int ix, iy, iz; cellTuple(gid, ix, iy, iz);
    Vec3 r = nodeCoord(ix, iy, iz, 0);
    r.x += dx / 2.; r.y += dy / 2.; r.z += dz / 2.;
    return r;
}

void GlobalFccGrid::cellNodes(int64_t gid, Vec3 out[14]) const
{
    // (di, dj, dk, basis) of the 14 points: corners 000,100,010,110,001,101,011,111, then the centres
    // of the +x,-x,+y,-y,+z,-z faces
    static const int off[14][4] = {
        {0,0,0,0}, {1,0,0,0}, {0,1,0,0}, {1,1,0,0}, {0,0,1,0}, {1,0,1,0}, {0,1,1,0}, {1,1,1,0},
        {1,0,0,1}, {0,0,0,1}, {0,1,0,2}, {0,0,0,2}, {0,0,1,3}, {0,0,0,3} };
    int ix, iy, iz; cellTuple(gid, ix, iy, iz);
    for (int p = 0; p < 14; ++p) out[p] = nodeCoord(ix + off[p][0], iy + off[p][1], iz + off[p][2], off[p][3]);
}

void GlobalFccGrid::faceNeighbors(int64_t gid, int64_t out[6]) const
{
    static const int step[6][3] = { {1,0,0}, {-1,0,0}, {0,1,0}, {0,-1,0}, {0,0,1}, {0,0,-1} };
    int ix, iy, iz; cellTuple(gid, ix, iy, iz);
    for (int f = 0; f < 6; ++f)
    {
        const int jx = std::min(std::max(0, ix + step[f][0]), nx - 1);
        const int jy = std::min(std::max(0, iy + step[f][1]), ny - 1);
        const int jz = std::min(std::max(0, iz + step[f][2]), nz - 1);
        out[f] = cellGid(jx, jy, jz);
    }
}

Vec3 cellPosition(const double* n)
{
    Vec3 c = { 0., 0., 0. };
    for (int p = 0; p < 14; ++p) { c.x += n[3 * p]; c.y += n[3 * p + 1]; c.z += n[3 * p + 2]; }
    const double inv = 1.0 / ((double)14);
    c.x *= inv; c.y *= inv; c.z *= inv;
    return c;
}

namespace {

// Domain centres.  Grid mode: src/initMC.cc:371-384 (ix outermost).  Random mode (xDom=yDom=zDom=0,
// single rank, 4 domains): src/initMC.cc:347-366 -- even-indexed cells picked with drand48(), which the
// reference never seeds, so the sequence is the libc default (erand48 below starts from glibc's
// actual unseeded state, X = 0, so every model built in a process sees the same sequence).  The reference
// builds the index triple as constructor arguments, which g++ evaluates right to left: z is drawn
// first, x last.
void domainCenters(const Parameters& params, const GlobalFccGrid& grid, int nCenters, std::vector<Vec3>& centers)
{
    const SimulationParameters& sp = params.simulationParams;
    if (sp.xDom == 0 && sp.yDom == 0 && sp.zDom == 0)
    {
        std::set<std::tuple<int, int, int> > picked;
        unsigned short state[3] = { 0, 0, 0 };      // glibc's unseeded drand48 starts from X = 0
        while ((int)centers.size() < nCenters)
        {
            const int iz = (int)(erand48(state) * (double)grid.nz / 2);
            const int iy = (int)(erand48(state) * (double)grid.ny / 2);
            const int ix = (int)(erand48(state) * (double)grid.nx / 2);
            if (!picked.insert(std::make_tuple(ix, iy, iz)).second) continue;
            centers.push_back(grid.cellCenter(grid.cellGid(2 * ix, 2 * iy, 2 * iz)));
        }
        return;
    }
    const double ddx = sp.lx / sp.xDom, ddy = sp.ly / sp.yDom, ddz = sp.lz / sp.zDom;
    for (int ix = 0; ix < sp.xDom; ++ix)
        for (int iy = 0; iy < sp.yDom; ++iy)
            for (int iz = 0; iz < sp.zDom; ++iz)
                centers.push_back(Vec3{ (0.5 + ix) * ddx, (0.5 + iy) * ddy, (0.5 + iz) * ddz });
}

int nearestCenter(const Vec3& r, const std::vector<Vec3>& centers)
{
    double best = 1e300; int who = -1;
    for (int i = 0; i < (int)centers.size(); ++i)
    {
        const double ex = r.x - centers[i].x, ey = r.y - centers[i].y, ez = r.z - centers[i].z;
        const double r2 = ex * ex + ey * ey + ez * ez;
        if (r2 < best) { best = r2; who = i; }      // strict <: an exact tie keeps the lower domain id
    }
    return who;
}

// normalised plane through three points (src/MC_Facet_Geometry.hh:18-39)
void facetPlane(const double* r0, const double* r1, const double* r2, double* out)
{
    double A = ((r1[1] - r0[1]) * (r2[2] - r0[2])) - ((r1[2] - r0[2]) * (r2[1] - r0[1]));
    double B = ((r1[2] - r0[2]) * (r2[0] - r0[0])) - ((r1[0] - r0[0]) * (r2[2] - r0[2]));
    double C = ((r1[0] - r0[0]) * (r2[1] - r0[1])) - ((r1[1] - r0[1]) * (r2[0] - r0[0]));
    double D = -1.0 * (A * r0[0] + B * r0[1] + C * r0[2]);
    double magnitude = std::sqrt(A * A + B * B + C * C);
    if (magnitude == 0.0) { A = 1.0; magnitude = 1.0; }
    const double inv = 1.0 / magnitude;
    out[0] = A * inv; out[1] = B * inv; out[2] = C * inv; out[3] = D * inv;
}

// src/MC_Domain.cc:334-355: centre = (sum of 14 points) / 14, volume = sum |a . (b x c)| / 6
double cellVolume(const double* n)
{
    double cx = 0., cy = 0., cz = 0.;
    for (int p = 0; p < 14; ++p) { cx += n[3 * p]; cy += n[3 * p + 1]; cz += n[3 * p + 2]; }
    cx /= 14; cy /= 14; cz /= 14;
    double volume = 0;
    for (int f = 0; f < 24; ++f)
    {
        const double* pa = n + 3 * kFacetPoints[f][0];
        const double* pb = n + 3 * kFacetPoints[f][1];
        const double* pc = n + 3 * kFacetPoints[f][2];
        const double ax = pa[0] - cx, ay = pa[1] - cy, az = pa[2] - cz;
        const double bx = pb[0] - cx, by = pb[1] - cy, bz = pb[2] - cz;
        const double qx = pc[0] - cx, qy = pc[1] - cy, qz = pc[2] - cz;
        const double crx = by * qz - bz * qy, cry = bz * qx - bx * qz, crz = bx * qy - by * qx;
        volume += std::abs(ax * crx + ay * cry + az * crz);
    }
    return volume / 6.0;
}

bool inside(const GeometryParameters& g, const Vec3& r)
{
    if (g.shape == GeometryParameters::BRICK)
        return (r.x >= g.xMin && r.x <= g.xMax) && (r.y >= g.yMin && r.y <= g.yMax) && (r.z >= g.zMin && r.z <= g.zMax);
    if (g.shape == GeometryParameters::SPHERE)
    {
        const double ex = r.x - g.xCenter, ey = r.y - g.yCenter, ez = r.z - g.zCenter;
        return std::sqrt(ex * ex + ey * ey + ez * ez) <= g.radius;
    }
    return false;
}

void boundaryEvents(const std::string& bc, uint8_t out[6])
{
    enum { Escape = 1, Reflect = 2 };
    if (bc == "reflect")      for (int f = 0; f < 6; ++f) out[f] = Reflect;
    else if (bc == "escape")  for (int f = 0; f < 6; ++f) out[f] = Escape;
    else if (bc == "octant")  for (int f = 0; f < 6; ++f) out[f] = (f % 2 == 0) ? Escape : Reflect;
    else throw std::runtime_error("boundaryCondition must be reflect | escape | octant, got '" + bc + "'");
}

} // namespace

void initMesh(const Parameters& params, const MaterialDatabase& db, int myRank, int nRanks,
              DecompositionInfo& ddc, std::vector<Domain>& domains)
{
    const SimulationParameters& sp = params.simulationParams;
    GlobalFccGrid grid(sp.nx, sp.ny, sp.nz, sp.lx, sp.ly, sp.lz);

    ddc.nRanks = nRanks; ddc.myRank = myRank;
    // src/initMC.cc:256-259: four randomly centred domains on a single rank without a domain grid, else one domain per rank.
    // Beyond the reference (which stops at `nRanks > 1 && nDomainsPerRank != 1`, src/initMC.cc:288-289): a domain grid with
    // k * nRanks centres gives every rank k consecutive domains (north_star: "one or more domains per GPU").
    const bool noGrid = sp.xDom == 0 && sp.yDom == 0 && sp.zDom == 0;
    const long long gridCenters = (long long)sp.xDom * sp.yDom * sp.zDom;
    if (noGrid) ddc.nDomainsPerRank = nRanks == 1 ? 4 : 1;
    else
    {
        if (gridCenters <= 0 || gridCenters % nRanks != 0)
            throw std::runtime_error("xDom*yDom*zDom must be a multiple of the number of ranks (GPUs)");
        ddc.nDomainsPerRank = (int)(gridCenters / nRanks);
    }
    const int nCenters = nRanks * ddc.nDomainsPerRank;
    domainCenters(params, grid, nCenters, ddc.centers);
    if ((int)ddc.centers.size() != nCenters)
        throw std::runtime_error("xDom*yDom*zDom must be a multiple of the number of ranks (GPUs)");
    for (int i = 0; i < ddc.nDomainsPerRank; ++i) ddc.myDomainGids.push_back(ddc.nDomainsPerRank * myRank + i);

    // owner domain and domain-local index of every global cell (local index = rank by ascending gid)
    const int64_t nGlobal = (int64_t)sp.nx * sp.ny * sp.nz;
    std::vector<int32_t> owner(nGlobal), localIndex(nGlobal);
    std::vector<int32_t> count(nCenters, 0);
    for (int64_t g = 0; g < nGlobal; ++g)
    {
        owner[g] = nCenters == 1 ? 0 : nearestCenter(grid.cellCenter(g), ddc.centers);
        localIndex[g] = count[owner[g]]++;
    }

    uint8_t bcEvent[6];
    boundaryEvents(sp.boundaryCondition, bcEvent);

    domains.clear();
    domains.resize(ddc.myDomainGids.size());
    for (size_t di = 0; di < domains.size(); ++di)
    {
        Domain& d = domains[di];
        d.globalDomain = ddc.myDomainGids[di];
        d.nCells = count[d.globalDomain];
        if (d.nCells == 0) throw std::runtime_error("a domain received no cells");
        const size_t n = d.nCells;
        d.cellGid.resize(n); d.nodes.resize(n * 42); d.planes.resize(n * 96);
        d.faceEvent.resize(n * 6); d.faceAdjCell.resize(n * 6); d.faceAdjDomain.resize(n * 6);
        d.faceNbrRank.resize(n * 6); d.faceAdjGlobalDomain.resize(n * 6);
        d.material.resize(n); d.volume.resize(n); d.cellId.resize(n); d.sourceTally.assign(n, 0);
    }

    auto materialAt = [&](const double* nodes) {
        const Vec3 where = cellPosition(nodes);
        std::string matName;                                   // last geometry containing the point wins
        for (const GeometryParameters& geom : params.geometryParams)
            if (inside(geom, where)) matName = geom.materialName;
        const int mat = db.findMaterial(matName);
        if (mat < 0) throw std::runtime_error("a mesh cell lies in no geometry / unknown material '" + matName + "'");
        return mat;
    };

    // Several ranks: every rank also evaluates volume * sourceRate of EVERY global cell, in global-id order, so that
    // the total source weight (src/MC_SourceNow.cc:41-57) can be summed in one canonical order on all ranks and an
    // N-rank run reproduces the single-rank run bit for bit instead of depending on an allreduce's rounding.
    ddc.globalVolumeRate.clear();
    if (nRanks > 1)
    {
        ddc.globalVolumeRate.resize(nGlobal);
        for (int64_t g = 0; g < nGlobal; ++g)
        {
            Vec3 pts[14];
            double nodes[42];
            grid.cellNodes(g, pts);
            for (int p = 0; p < 14; ++p) { nodes[3 * p] = pts[p].x; nodes[3 * p + 1] = pts[p].y; nodes[3 * p + 2] = pts[p].z; }
            ddc.globalVolumeRate[g] = cellVolume(nodes) * db.mat[materialAt(nodes)].sourceRate;
        }
    }

    for (int64_t g = 0; g < nGlobal; ++g)
    {
        if (ddc.rankOf(owner[g]) != myRank) continue;
        Domain& d = domains[ddc.indexOf(owner[g])];
        const int c = localIndex[g];
        d.cellGid[c] = (int32_t)g;

        Vec3 pts[14];
        grid.cellNodes(g, pts);
        double* nodes = &d.nodes[(size_t)c * 42];
        for (int p = 0; p < 14; ++p) { nodes[3 * p] = pts[p].x; nodes[3 * p + 1] = pts[p].y; nodes[3 * p + 2] = pts[p].z; }
        for (int f = 0; f < 24; ++f)
            facetPlane(nodes + 3 * kFacetPoints[f][0], nodes + 3 * kFacetPoints[f][1], nodes + 3 * kFacetPoints[f][2],
                       &d.planes[((size_t)c * 24 + f) * 4]);

        int64_t nbr[6];
        grid.faceNeighbors(g, nbr);
        for (int f = 0; f < 6; ++f)
        {
            const size_t k = (size_t)c * 6 + f;
            const int nbrDomain = owner[nbr[f]];
            d.faceAdjGlobalDomain[k] = nbrDomain;
            d.faceAdjDomain[k] = ddc.indexOf(nbrDomain);
            d.faceAdjCell[k] = localIndex[nbr[f]];
            d.faceNbrRank[k] = -1;
            if (nbr[f] == g)                                   d.faceEvent[k] = bcEvent[f];
            else if (ddc.rankOf(nbrDomain) == myRank)          d.faceEvent[k] = 3;   // Transit_On_Processor
            else { d.faceEvent[k] = 4; d.faceNbrRank[k] = ddc.rankOf(nbrDomain); }   // Transit_Off_Processor
        }

        d.volume[c] = cellVolume(nodes);
        d.material[c] = materialAt(nodes);

        Vec3 centre = { 0., 0., 0. };                          // src/MC_Domain.cc:321-330 (divide, not multiply)
        for (int p = 0; p < 14; ++p) { centre.x += pts[p].x; centre.y += pts[p].y; centre.z += pts[p].z; }
        centre.x /= 14; centre.y /= 14; centre.z /= 14;
        d.cellId[c] = (uint64_t)grid.whichCell(centre) * UINT64_C(0x0100000000);
    }
}

} // namespace qsb
