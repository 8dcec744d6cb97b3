// TEST INFRASTRUCTURE ONLY -- never linked into or called from the product.
//
// qs_dump: drives the UNMODIFIED reference (compiled from /root/reference/src where
// it lies, see oracle/Makefile) through its own public functions and writes what the
// reference holds in memory to a directory of "QSD1" array files, so that the CPU
// restatement (oracle/qs_oracle.c) and the host model (quicksilver_b200/csrc) can be
// pinned bit-for-bit against the real thing:
//   problem.qsd        mesh (nodes, planes, adjacency), cell state, nuclear data, materials
//   cycle_NNN.qsd      processing vaults after cycleInit (= tracking input), processed
//                      vaults after cycleTracking (= census), balance counters, scalar flux
//
// The reference's main.cc is included as-is (its main renamed) so that its file-local
// cycleInit / cycleTracking / cycleFinalize (src/main.cc:96,138,310) are the code that runs.
//
// usage: QS_DUMP_DIR=out [QS_DUMP_PARTICLE_CYCLES=K] qs_dump <same CLI as qs>
#define main qs_reference_main
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

namespace {

struct QsdWriter
{
   FILE* f;
   explicit QsdWriter(const std::string& path) : f(fopen(path.c_str(), "wb"))
   {
      if (!f) { perror(path.c_str()); exit(3); }
      fwrite("QSD1", 1, 4, f);
   }
   ~QsdWriter() { fclose(f); }
   // dtype: 'd' f64, 'i' i32, 'u' u64, 'b' raw bytes
   void put(const char* name, char dtype, const void* data, uint64_t count, uint64_t inner = 1)
   {
      uint32_t nlen = strlen(name);
      uint64_t esz = dtype == 'd' ? 8 : dtype == 'i' ? 4 : dtype == 'u' ? 8 : 1;
      fwrite(&nlen, 4, 1, f); fwrite(name, 1, nlen, f);
      fwrite(&dtype, 1, 1, f);
      fwrite(&count, 8, 1, f); fwrite(&inner, 8, 1, f);
      if (count * inner) fwrite(data, esz, count * inner, f);
   }
};

std::string dumpDir() { const char* d = getenv("QS_DUMP_DIR"); return d ? d : "qs_dump_out"; }

void dumpProblem()
{
   QsdWriter w(dumpDir() + "/problem.qsd");
   const Parameters& pp = mcco->_params;
   int nDomains = mcco->domain.size();
   int nGroups = mcco->_nuclearData->_numEnergyGroups;
   std::vector<int> hdr = { nDomains, nGroups, (int)mcco->_nuclearData->_isotopes.size(),
                            (int)mcco->_materialDatabase->_mat.size(),
                            pp.simulationParams.nx, pp.simulationParams.ny, pp.simulationParams.nz };
   w.put("header", 'i', hdr.data(), hdr.size());
   w.put("energies", 'd', &mcco->_nuclearData->_energies[0], mcco->_nuclearData->_energies.size());

   // nuclear data: [iso][react] type, nuBar, sigma[g]
   {
      std::vector<int> nReact, rtype; std::vector<double> nuBar, sigma;
      for (int i = 0; i < mcco->_nuclearData->_isotopes.size(); ++i)
      {
         auto& rx = mcco->_nuclearData->_isotopes[i]._species[0]._reactions;
         nReact.push_back(rx.size());
         for (int r = 0; r < rx.size(); ++r)
         {
            rtype.push_back((int)rx[r]._reactionType);
            nuBar.push_back(rx[r]._nuBar);
            for (int g = 0; g < nGroups; ++g) sigma.push_back(rx[r]._crossSection[g]);
         }
      }
      w.put("iso_nreact", 'i', nReact.data(), nReact.size());
      w.put("react_type", 'i', rtype.data(), rtype.size());
      w.put("react_nubar", 'd', nuBar.data(), nuBar.size());
      w.put("react_sigma", 'd', sigma.data(), rtype.size(), nGroups);
   }
   // materials
   {
      std::string names; std::vector<double> mass, af; std::vector<int> nIso, gid;
      for (int m = 0; m < mcco->_materialDatabase->_mat.size(); ++m)
      {
         auto& mat = mcco->_materialDatabase->_mat[m];
         names += mat._name; names += '\n';
         mass.push_back(mat._mass);
         nIso.push_back(mat._iso.size());
         for (int k = 0; k < mat._iso.size(); ++k) { gid.push_back(mat._iso[k]._gid); af.push_back(mat._iso[k]._atomFraction); }
      }
      w.put("mat_names", 'b', names.data(), names.size());
      w.put("mat_mass", 'd', mass.data(), mass.size());
      w.put("mat_niso", 'i', nIso.data(), nIso.size());
      w.put("mat_iso_gid", 'i', gid.data(), gid.size());
      w.put("mat_iso_af", 'd', af.data(), af.size());
   }
   for (int d = 0; d < nDomains; ++d)
   {
      MC_Domain& dom = mcco->domain[d];
      int nCells = dom.cell_state.size();
      char name[64];
      std::vector<int> info = { dom.global_domain, nCells, (int)dom.mesh._node.size(), (int)dom.mesh._nbrRank.size() };
      snprintf(name, 64, "d%d_info", d);      w.put(name, 'i', info.data(), info.size());
      snprintf(name, 64, "d%d_nbr_rank", d);  w.put(name, 'i', &dom.mesh._nbrRank[0], dom.mesh._nbrRank.size());
      snprintf(name, 64, "d%d_nbr_gid", d);   w.put(name, 'i', &dom.mesh._nbrDomainGid[0], dom.mesh._nbrDomainGid.size());
      std::vector<double> nodes(nCells * 14 * 3), planes(nCells * 24 * 4), vol(nCells), dens(nCells);
      std::vector<int> adj(nCells * 24 * 8), mat(nCells), fpts(nCells * 24 * 3);
      std::vector<uint64_t> ids(nCells);
      for (int c = 0; c < nCells; ++c)
      {
         const MC_Facet_Adjacency_Cell& cc = dom.mesh._cellConnectivity[c];
         for (int p = 0; p < 14; ++p)
         {
            const MC_Vector& v = dom.mesh._node[cc._point[p]];
            nodes[(c * 14 + p) * 3 + 0] = v.x; nodes[(c * 14 + p) * 3 + 1] = v.y; nodes[(c * 14 + p) * 3 + 2] = v.z;
         }
         for (int f = 0; f < 24; ++f)
         {
            const MC_General_Plane& pl = dom.mesh._cellGeometry[c]._facet[f];
            double* q = &planes[(c * 24 + f) * 4]; q[0] = pl.A; q[1] = pl.B; q[2] = pl.C; q[3] = pl.D;
            const Subfacet_Adjacency& s = cc._facet[f].subfacet;
            int* a = &adj[(c * 24 + f) * 8];
            a[0] = (int)s.event; a[1] = s.adjacent.domain; a[2] = s.adjacent.cell; a[3] = s.adjacent.facet;
            a[4] = s.neighbor_index; a[5] = s.neighbor_global_domain; a[6] = s.neighbor_foreman; a[7] = s.current.facet;
            // facet points as positions in the cell's own 14-point list
            for (int k = 0; k < 3; ++k)
            {
               int pid = cc._facet[f].point[k], local = -1;
               for (int p = 0; p < 14; ++p) if (cc._point[p] == pid) local = p;
               fpts[(c * 24 + f) * 3 + k] = local;
            }
         }
         vol[c] = dom.cell_state[c]._volume; dens[c] = dom.cell_state[c]._cellNumberDensity;
         mat[c] = dom.cell_state[c]._material; ids[c] = dom.cell_state[c]._id;
      }
      snprintf(name, 64, "d%d_nodes", d);   w.put(name, 'd', nodes.data(), nCells, 42);
      snprintf(name, 64, "d%d_planes", d);  w.put(name, 'd', planes.data(), nCells, 96);
      snprintf(name, 64, "d%d_adj", d);     w.put(name, 'i', adj.data(), nCells, 192);
      snprintf(name, 64, "d%d_fpts", d);    w.put(name, 'i', fpts.data(), nCells, 72);
      snprintf(name, 64, "d%d_volume", d);  w.put(name, 'd', vol.data(), nCells);
      snprintf(name, 64, "d%d_density", d); w.put(name, 'd', dens.data(), nCells);
      snprintf(name, 64, "d%d_material", d);w.put(name, 'i', mat.data(), nCells);
      snprintf(name, 64, "d%d_cell_id", d); w.put(name, 'u', ids.data(), nCells);
   }
}

void gather(bool processing, std::vector<MC_Base_Particle>& out)
{
   ParticleVaultContainer& pvc = *mcco->_particleVaultContainer;
   uint64_t nv = processing ? pvc.processingSize() : pvc.processedSize();
   for (uint64_t v = 0; v < nv; ++v)
   {
      ParticleVault* vault = processing ? pvc.getTaskProcessingVault(v) : pvc.getTaskProcessedVault(v);
      for (size_t j = 0; j < vault->size(); ++j) out.push_back((*vault)[j]);
   }
}

void balanceArray(uint64_t t[13])
{
   Balance& b = mcco->_tallies->_balanceTask[0];   // order of Tallies.cc:31-43
   uint64_t v[13] = { b._absorb, b._census, b._escape, b._collision, b._end, b._fission, b._produce,
                      b._scatter, b._start, b._source, b._rr, b._split, b._numSegments };
   memcpy(t, v, sizeof(v));
}

} // namespace

int main(int argc, char** argv)
{
   static_assert(sizeof(MC_Base_Particle) == 136, "MC_Base_Particle layout");
   mpiInit(&argc, &argv);
   Parameters params = getParameters(argc, argv);
   mcco = initMC(params);
   int loadBalance = params.simulationParams.loadBalance;
   const char* pc = getenv("QS_DUMP_PARTICLE_CYCLES");
   int particleCycles = pc ? atoi(pc) : 1;   // dump vaults for the first K cycles, tallies for all
   std::string mk = "mkdir -p " + dumpDir(); if (system(mk.c_str())) return 3;
   dumpProblem();

   const int nSteps = params.simulationParams.nSteps;
   for (int ii = 0; ii < nSteps; ++ii)
   {
      char fn[64]; snprintf(fn, 64, "/cycle_%03d.qsd", ii);
      QsdWriter w(dumpDir() + fn);

      cycleInit(bool(loadBalance));
      double spw = mcco->source_particle_weight;
      w.put("source_particle_weight", 'd', &spw, 1);
      uint64_t bal[13]; balanceArray(bal);
      w.put("balance_after_init", 'u', bal, 13);
      if (ii < particleCycles)
      {
         std::vector<MC_Base_Particle> in; gather(true, in);
         w.put("tracking_input", 'b', in.data(), in.size(), 136);
      }

      cycleTracking(mcco);

      mcco->_tallies->SumTasks();
      balanceArray(bal);
      bal[4] = mcco->_particleVaultContainer->sizeProcessed();   // _end, as cycleFinalize sets it
      w.put("balance", 'u', bal, 13);
      for (int d = 0; d < mcco->domain.size(); ++d)
      {
         ScalarFluxTask& t = mcco->_tallies->_scalarFluxDomain[d]._task[0];
         int nCells = t._cell.size(), nGroups = t._cell[0].size();
         std::vector<double> flux((size_t)nCells * nGroups, 0.0);
         for (int rep = 0; rep < mcco->_tallies->GetNumFluxReplications(); ++rep)
            for (int c = 0; c < nCells; ++c)
               for (int g = 0; g < nGroups; ++g)
                  flux[(size_t)c * nGroups + g] += mcco->_tallies->_scalarFluxDomain[d]._task[rep]._cell[c]._group[g];
         char name[64]; snprintf(name, 64, "d%d_flux", d);
         if (ii < particleCycles) w.put(name, 'd', flux.data(), nCells, nGroups);
      }
      double fsum = mcco->_tallies->ScalarFluxSum(mcco);
      w.put("scalar_flux_sum", 'd', &fsum, 1);
      if (ii < particleCycles)
      {
         std::vector<MC_Base_Particle> out; gather(false, out);
         w.put("census", 'b', out.data(), out.size(), 136);
      }

      cycleFinalize();
      mcco->fast_timer->Last_Cycle_Report(0, 0, 1, mcco->processor_info->comm_mc_world);
   }
   gameOver();
   coralBenchmarkCorrectness(mcco, params);
   mpiFinalize();
   return 0;
}
