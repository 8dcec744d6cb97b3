// QsbBinding.cc -- the REFERENCE-SIDE binding of libqsb.so: what a Quicksilver maintainer adds to run cycle tracking on the
// B200 library while everything else stays the reference's own code.  It is compiled against the unmodified reference
// sources where they lie (oracle/Makefile: `make qsb_binding` -> oracle/_ref/qs_qsb) and is the source INTEGRATION.md
// section A is generated from -- so the drop-in really has been dropped in.
//
//   reference code that runs unchanged : mpiInit, getParameters, printParameters, initMC (mesh, nuclear data, materials),
//                                        cycleInit (MC_SourceNow, PopulationControl, RouletteLowWeightParticles),
//                                        cycleFinalize (Tallies::CycleFinalize, EnergySpectrum, Fluence), the MC_Fast_Timer
//                                        table + Figure Of Merit, coralBenchmarkCorrectness
//   replaced                           : the body of cycleTracking(MonteCarlo*), src/main.cc:138-307 (the per-vault kernel loop
//                                        that ends in CycleTrackingGuts, src/CycleTracking.hh:8-14) -> cycleTrackingQsb below
//
// The reference's main.cc is included as it is with its main() renamed, so that its file-local cycleInit / cycleFinalize /
// gameOver (src/main.cc:87-121,310-324) are the functions that run; the new main() below is src/main.cc:38-85 with ONE line
// changed (cycleTracking -> cycleTrackingQsb) and the attach call after initMC.  Nothing of this repository's host model
// (qsb_mc_*) is used: only the device context (qsb_create / qsb_cycle_begin / qsb_put_particles / qsb_track / qsb_get_*).
//
// Single rank (the reference build here uses its serial MPI stubs).  QSB_FAST=1 selects the fast kernels; the default is
// the bit-exact validation build, with which `qs_qsb -i CTS2_1.inp` prints the reference's own cycle table.
#define main(...) qs_reference_main(__VA_ARGS__)      // function-like: MC_Fast_Timer::main (an enumerator) is left alone
#include "main.cc"
#undef main

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "MC_Domain.hh"
#include "NuclearData.hh"
#include "MaterialDatabase.hh"
#include "MC_Base_Particle.hh"
#include "qsb.h"

namespace {

qsb_ctx* g_ctx = nullptr;

// storage behind the qsb_image pointers (the library copies what it needs in qsb_create; kept for the run anyway)
struct ImageStorage
{
   std::vector<int32_t>  domainCellOffset, domainGid, cellGid, cellMaterial, faceAdjCell, faceAdjDomain, faceNbrRank, matNIso, matNReact;
   std::vector<double>   planes, nodes, cellVolume, energies, matMass, matNuBar, xsTotal, xsReact;
   std::vector<uint64_t> cellId;
   std::vector<uint8_t>  faceEvent, matReactType, matPeriodic;
} g_store;

// Flatten what the tracking path reads of the reference's MonteCarlo object into a qsb_image (include/qsb.h): all the
// rank's domains concatenated into one cell index space, geometry per cell, adjacency per FACE (the four facets of a face
// share event and neighbour, src/MC_Domain.cc:292-318), the multigroup tables per material.
void qsbAttach(MonteCarlo* mc, const Parameters& params)
{
   static_assert(sizeof(MC_Base_Particle) == sizeof(qsb_base_particle), "MC_Base_Particle must be the 136-byte record of qsb.h");
   ImageStorage& s = g_store;
   const int nDomains = (int)mc->domain.size();
   const int nGroups = mc->_nuclearData->_numEnergyGroups;
   const int nMat = (int)mc->_materialDatabase->_mat.size();

   s.domainCellOffset.assign(nDomains + 1, 0);
   for (int d = 0; d < nDomains; ++d)
   {
      s.domainCellOffset[d + 1] = s.domainCellOffset[d] + (int)mc->domain[d].cell_state.size();
      s.domainGid.push_back(mc->domain[d].global_domain);
   }
   for (int d = 0; d < nDomains; ++d)
   {
      MC_Domain& dom = mc->domain[d];
      const int nCells = (int)dom.cell_state.size();
      for (int c = 0; c < nCells; ++c)
      {
         const MC_Facet_Adjacency_Cell& cc = dom.mesh._cellConnectivity[c];
         for (int p = 0; p < 14; ++p)                                                  // the cell's own 14 points
         {
            const MC_Vector& v = dom.mesh._node[cc._point[p]];
            s.nodes.push_back(v.x); s.nodes.push_back(v.y); s.nodes.push_back(v.z);
         }
         for (int f = 0; f < 24; ++f)                                                  // facet planes A, B, C, D
         {
            const MC_General_Plane& pl = dom.mesh._cellGeometry[c]._facet[f];
            s.planes.push_back(pl.A); s.planes.push_back(pl.B); s.planes.push_back(pl.C); s.planes.push_back(pl.D);
         }
         for (int face = 0; face < 6; ++face)
         {
            const Subfacet_Adjacency& sub = cc._facet[4 * face].subfacet;
            const int event = (int)sub.event;
            s.faceEvent.push_back((uint8_t)event);
            if (event == QSB_ADJ_TRANSIT_ON)
            {
               s.faceAdjCell.push_back(s.domainCellOffset[sub.adjacent.domain] + sub.adjacent.cell);
               s.faceAdjDomain.push_back(sub.adjacent.domain);
               s.faceNbrRank.push_back(-1);
            }
            else if (event == QSB_ADJ_TRANSIT_OFF)
            {
               s.faceAdjCell.push_back(sub.adjacent.cell);                              // local to the neighbour's domain
               s.faceAdjDomain.push_back(sub.adjacent.domain);
               s.faceNbrRank.push_back(sub.neighbor_foreman);
            }
            else                                                                        // escape / reflection: the cell itself
            {
               s.faceAdjCell.push_back(s.domainCellOffset[d] + c);
               s.faceAdjDomain.push_back(d);
               s.faceNbrRank.push_back(-1);
            }
         }
         s.cellVolume.push_back(dom.cell_state[c]._volume);
         s.cellMaterial.push_back(dom.cell_state[c]._material);
         s.cellId.push_back(dom.cell_state[c]._id);
         s.cellGid.push_back((int32_t)(dom.cell_state[c]._id >> 32));                   // src/MC_Domain.cc:390: id = global cell << 32
      }
   }

   // nuclear data: per material, the (isotope, reaction) table in the scan order of CollisionEvent (src/CollisionEvent.cc:67-83)
   s.energies.assign(&mc->_nuclearData->_energies[0], &mc->_nuclearData->_energies[0] + mc->_nuclearData->_energies.size());
   int maxReact = 1, nIsoTotal = 0;
   for (int m = 0; m < nMat; ++m)
   {
      int n = 0;
      auto& mat = mc->_materialDatabase->_mat[m];
      for (int i = 0; i < (int)mat._iso.size(); ++i) n += (int)mc->_nuclearData->_isotopes[mat._iso[i]._gid]._species[0]._reactions.size();
      if (n > maxReact) maxReact = n;
      nIsoTotal += (int)mat._iso.size();
   }
   s.xsTotal.assign((size_t)nMat * nGroups, 0.0);
   s.xsReact.assign((size_t)nMat * nGroups * maxReact, 0.0);
   s.matReactType.assign((size_t)nMat * maxReact, QSB_REACT_UNDEFINED);
   for (int m = 0; m < nMat; ++m)
   {
      auto& mat = mc->_materialDatabase->_mat[m];
      const int nIso = (int)mat._iso.size();
      const int nReact0 = nIso ? (int)mc->_nuclearData->_isotopes[mat._iso[0]._gid]._species[0]._reactions.size() : 0;
      s.matNIso.push_back(nIso);
      s.matNReact.push_back(nReact0);
      s.matMass.push_back(mat._mass);
      double nuBar = 0.0;
      bool periodic = true;
      const double cellNumberDensity = 1.0;                                            // src/MC_Domain.cc:387
      for (int g = 0; g < nGroups; ++g)
      {
         double sum = 0.0;                                                              // weightedMacroscopicCrossSection, :59-80
         int k = 0;
         for (int i = 0; i < nIso; ++i)
         {
            const int gid = mat._iso[i]._gid;
            const double af = mat._iso[i]._atomFraction;
            auto& reactions = mc->_nuclearData->_isotopes[gid]._species[0]._reactions;
            if (af == 0.0 || cellNumberDensity == 0.0) sum += 1e-20;                    // macroscopicCrossSection, :31
            else sum += af * cellNumberDensity * mc->_nuclearData->getTotalCrossSection(gid, g);
            for (int r = 0; r < (int)reactions.size(); ++r, ++k)
            {
               const double v = (af == 0.0 || cellNumberDensity == 0.0) ? 1e-20 : af * cellNumberDensity * reactions[r]._crossSection[g];
               s.xsReact[((size_t)m * nGroups + g) * maxReact + k] = v;
               if (g == 0)
               {
                  s.matReactType[(size_t)m * maxReact + k] = (uint8_t)reactions[r]._reactionType;
                  if (reactions[r]._reactionType == NuclearDataReaction::Fission) nuBar = reactions[r]._nuBar;
               }
               // "periodic": every isotope of the material carries isotope 0's reaction table (true for every deck the
               // reference's grammar can express); lets the kernel select a reaction with one division
               if ((int)reactions.size() != nReact0 || r >= nReact0 ||
                   memcmp(&v, &s.xsReact[((size_t)m * nGroups + g) * maxReact + (r < nReact0 ? r : 0)], 8) != 0 ||
                   reactions[r]._reactionType != mc->_nuclearData->_isotopes[mat._iso[0]._gid]._species[0]._reactions[r < nReact0 ? r : 0]._reactionType)
                  periodic = false;
            }
         }
         s.xsTotal[(size_t)m * nGroups + g] = sum;
      }
      s.matNuBar.push_back(nuBar);
      s.matPeriodic.push_back(periodic ? 1 : 0);
   }

   qsb_image im;
   memset(&im, 0, sizeof im);
   im.abi_version = QSB_ABI_VERSION;
   im.n_domains = nDomains; im.n_cells = s.domainCellOffset[nDomains]; im.n_groups = nGroups; im.n_materials = nMat;
   im.n_isotopes = nIsoTotal; im.max_reactions_per_material = maxReact;
   im.my_rank = mc->processor_info->rank; im.n_ranks = mc->processor_info->num_processors;
   im.global_nx = params.simulationParams.nx; im.global_ny = params.simulationParams.ny; im.global_nz = params.simulationParams.nz;
   im.global_lx = params.simulationParams.lx; im.global_ly = params.simulationParams.ly; im.global_lz = params.simulationParams.lz;
   im.domain_cell_offset = s.domainCellOffset.data(); im.domain_gid = s.domainGid.data();
   im.planes = s.planes.data(); im.nodes = s.nodes.data(); im.cell_gid = s.cellGid.data();
   im.cell_material = s.cellMaterial.data(); im.cell_volume = s.cellVolume.data(); im.cell_id = s.cellId.data();
   im.face_event = s.faceEvent.data(); im.face_adj_cell = s.faceAdjCell.data();
   im.face_adj_domain = s.faceAdjDomain.data(); im.face_nbr_rank = s.faceNbrRank.data();
   im.energies = s.energies.data(); im.mat_n_isotopes = s.matNIso.data(); im.mat_n_reactions = s.matNReact.data();
   im.mat_mass = s.matMass.data(); im.mat_nu_bar = s.matNuBar.data(); im.mat_react_type = s.matReactType.data();
   im.xs_total = s.xsTotal.data(); im.xs_react = s.xsReact.data(); im.mat_periodic = s.matPeriodic.data();

   if (const char* dir = getenv("QSB_BINDING_IMAGE_DUMP"))                              // tests: the image as raw arrays, one file each
   {
      auto dump = [&](const char* name, const void* data, size_t bytes) {
         const std::string path = std::string(dir) + "/" + name + ".bin";
         FILE* f = fopen(path.c_str(), "wb");
         if (!f) { perror(path.c_str()); exit(5); }
         if (bytes) fwrite(data, 1, bytes, f);
         fclose(f);
      };
      const int32_t hdr[12] = { im.n_domains, im.n_cells, im.n_groups, im.n_materials, im.n_isotopes, im.max_reactions_per_material,
                                im.my_rank, im.n_ranks, im.global_nx, im.global_ny, im.global_nz, 0 };
      dump("header", hdr, sizeof hdr);
#define QSB_DUMP(v_) dump(#v_, s.v_.data(), s.v_.size() * sizeof(s.v_[0]))
      QSB_DUMP(domainCellOffset); QSB_DUMP(domainGid); QSB_DUMP(planes); QSB_DUMP(nodes); QSB_DUMP(cellGid); QSB_DUMP(cellMaterial);
      QSB_DUMP(cellVolume); QSB_DUMP(cellId); QSB_DUMP(faceEvent); QSB_DUMP(faceAdjCell); QSB_DUMP(faceAdjDomain); QSB_DUMP(faceNbrRank);
      QSB_DUMP(energies); QSB_DUMP(matNIso); QSB_DUMP(matNReact); QSB_DUMP(matMass); QSB_DUMP(matNuBar); QSB_DUMP(matReactType);
      QSB_DUMP(xsTotal); QSB_DUMP(xsReact); QSB_DUMP(matPeriodic);
#undef QSB_DUMP
   }

   qsb_options opt;
   memset(&opt, 0, sizeof opt);
   const char* fast = getenv("QSB_FAST");
   opt.validation = (fast && atoi(fast) != 0) ? 0 : 1;                                  // default: the bit-exact build
   // room for a cycle's secondaries on top of the population the deck asks for (fission: nuBar <= 3 per collision)
   opt.particle_capacity = 8ull * (uint64_t)params.simulationParams.nParticles + (1ull << 16);
   const int rc = qsb_create(/*device*/ 0, &im, mc->time_info->time_step, &opt, &g_ctx);
   if (rc != QSB_OK)
   {
      fprintf(stderr, "qsb_create failed (%d): %s\n", rc, qsb_last_error(nullptr));      // no CPU fallback: the run ends here
      exit(3);
   }
}

void qsbCheck(int rc, const char* what)
{
   if (rc == QSB_OK) return;
   fprintf(stderr, "%s failed (%d): %s\n", what, rc, qsb_last_error(g_ctx));
   exit(4);
}

// Replaces the body of cycleTracking(MonteCarlo*), src/main.cc:138-307, for one rank: the processing vaults go to the
// device as they are (136-byte records), ONE call tracks every history of the cycle to its end -- secondaries included,
// which the reference feeds back through the extra vaults and further kernel launches -- and the census and the tallies
// come back into the reference's own containers, so that cycleFinalize and the next cycleInit are none the wiser.
void cycleTrackingQsb(MonteCarlo* mc)
{
   MC_FASTTIMER_START(MC_Fast_Timer::cycleTracking);
   ParticleVaultContainer& pvc = *mc->_particleVaultContainer;

   MC_FASTTIMER_START(MC_Fast_Timer::cycleTracking_Kernel);
   qsbCheck(qsb_cycle_begin(g_ctx, 0), "qsb_cycle_begin");
   for (uint64_t v = 0; v < pvc.processingSize(); ++v)                                  // host vaults in
   {
      ParticleVault* pv = pvc.getTaskProcessingVault(v);
      if (pv->size())
         qsbCheck(qsb_put_particles(g_ctx, reinterpret_cast<const qsb_base_particle*>(&(*pv)[0]), pv->size()), "qsb_put_particles");
      pv->clear();
   }
   qsb_track_stats stats;
   qsbCheck(qsb_track(g_ctx, &stats), "qsb_track");
   MC_FASTTIMER_STOP(MC_Fast_Timer::cycleTracking_Kernel);

   MC_FASTTIMER_START(MC_Fast_Timer::cycleTracking_MPI);
   // census out -> processed vaults, each filled to the container's vault size (src/ParticleVaultContainer.cc:111-128)
   std::vector<qsb_base_particle> census(stats.n_census);
   uint64_t n = 0;
   qsbCheck(qsb_get_census(g_ctx, census.data(), census.size(), &n), "qsb_get_census");
   const uint64_t vaultSize = pvc.getVaultSize();
   for (uint64_t i = 0; i < n;)
   {
      ParticleVault* out = pvc.getTaskProcessedVault(pvc.getFirstEmptyProcessedVault());
      for (uint64_t k = 0; k < vaultSize && i < n; ++k, ++i)
      {
         MC_Base_Particle p;
         memcpy(&p, &census[i], sizeof p);
         out->pushBaseParticle(p);
      }
   }
   // tallies out: the eight tracking counters of Balance (src/Tallies.hh:36-100) and the scalar flux (src/Tallies.hh:351-354)
   uint64_t b[QSB_BAL_COUNT];
   qsbCheck(qsb_get_balance(g_ctx, b), "qsb_get_balance");
   Balance& t = mc->_tallies->_balanceTask[0];
   t._absorb += b[QSB_BAL_ABSORB];       t._census += b[QSB_BAL_CENSUS];   t._escape += b[QSB_BAL_ESCAPE];
   t._collision += b[QSB_BAL_COLLISION]; t._fission += b[QSB_BAL_FISSION]; t._produce += b[QSB_BAL_PRODUCE];
   t._scatter += b[QSB_BAL_SCATTER];     t._numSegments += b[QSB_BAL_NUM_SEGMENTS];
   const int nGroups = mc->_nuclearData->_numEnergyGroups;
   std::vector<double> flux((size_t)g_store.domainCellOffset.back() * nGroups);
   qsbCheck(qsb_get_scalar_flux(g_ctx, flux.data()), "qsb_get_scalar_flux");
   for (int d = 0; d < (int)mc->domain.size(); ++d)
   {
      ScalarFluxTask& task = mc->_tallies->_scalarFluxDomain[d]._task[0];
      const size_t first = (size_t)g_store.domainCellOffset[d];
      for (size_t c = 0; c < task._cell.size(); ++c)
         for (int g = 0; g < nGroups; ++g)
            task._cell[c]._group[g] += flux[(first + c) * nGroups + g];
   }
   MC_FASTTIMER_STOP(MC_Fast_Timer::cycleTracking_MPI);
   MC_FASTTIMER_STOP(MC_Fast_Timer::cycleTracking);
}

} // namespace

// src/main.cc:38-85 with the tracking call replaced
int main(int argc, char** argv)
{
   mpiInit(&argc, &argv);
   printBanner(GIT_VERS, GIT_HASH);

   Parameters params = getParameters(argc, argv);
   printParameters(params, cout);

   mcco = initMC(params);
   qsbAttach(mcco, params);                       // <- added: hand the problem to the device library, once

   int loadBalance = params.simulationParams.loadBalance;

   MC_FASTTIMER_START(MC_Fast_Timer::main);

   const int nSteps = params.simulationParams.nSteps;

   for (int ii = 0; ii < nSteps; ++ii)
   {
      cycleInit(bool(loadBalance));
      cycleTrackingQsb(mcco);                     // <- was: cycleTracking(mcco)
      cycleFinalize();

      mcco->fast_timer->Last_Cycle_Report(
            params.simulationParams.cycleTimers,
            mcco->processor_info->rank,
            mcco->processor_info->num_processors,
            mcco->processor_info->comm_mc_world);
   }

   MC_FASTTIMER_STOP(MC_Fast_Timer::main);

   gameOver();

   coralBenchmarkCorrectness(mcco, params);

   qsb_destroy(g_ctx);
   delete mcco;

   mpiFinalize();
   return 0;
}
