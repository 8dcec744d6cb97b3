// NuclearData.hh -- multigroup cross-section tables and materials of the host model.
//
// What the reference keeps as NuclearData/_isotopes[i]._species[0]._reactions[r]._crossSection[g]
// plus MaterialDatabase (src/NuclearData.hh:55-120, src/MaterialDatabase.hh) is held here as flat,
// contiguous tables, because the only consumer is the flattened device image:
//   sigma[iso][react][group]   microscopic cross sections (bit-identical to the reference's)
//   per material: isotope gid range, atom fraction, mass, nuBar, reaction types
#ifndef QSB_NUCLEAR_DATA_HH
#define QSB_NUCLEAR_DATA_HH

#include <cstdint>
#include <string>
#include <vector>
#include "Parameters.hh"

namespace qsb {

struct ReactionType { enum Enum { Undefined = 0, Scatter = 1, Absorption = 2, Fission = 3 }; };

struct Isotope
{
    int nReactions = 0;
    size_t sigmaOffset = 0;                 // into NuclearData::sigma, [react][group]
    std::vector<uint8_t> reactionType;      // [react]
    double nuBar = 0;
};

struct Material
{
    std::string name;
    double mass = 1000.0;
    double sourceRate = 0.0;
    double nuBar = 0.0;
    std::vector<int> isoGid;                // index into NuclearData::isotopes
    std::vector<double> atomFraction;
};

class NuclearData
{
public:
    NuclearData(int numGroups, double energyLow, double energyHigh);

    // nReactions reactions cycling scatter / fission / absorption, each scaled so that the group
    // containing 1 MeV carries its share of totalCrossSection (src/NuclearData.cc:12-42,122-187).
    int addIsotope(int nReactions, const CrossSectionParameters& fission, const CrossSectionParameters& scatter,
                   const CrossSectionParameters& absorption, double nuBar, double totalCrossSection,
                   double fissionWeight, double scatterWeight, double absorptionWeight);

    int getEnergyGroup(double energy) const;                 // src/NuclearData.cc:208-227
    double sigmaOf(int iso, int react, int group) const { return sigma[isotopes[iso].sigmaOffset + (size_t)react * numGroups + group]; }
    double totalCrossSection(int iso, int group) const;      // src/NuclearData.cc:231-242

    int numGroups;
    std::vector<double> energies;       // numGroups + 1 edges
    std::vector<Isotope> isotopes;
    std::vector<double> sigma;
};

struct MaterialDatabase
{
    std::vector<Material> mat;
    int findMaterial(const std::string& name) const;
};

// materials in name order, each with nIsotopes private isotopes of atom fraction 1/nIsotopes
// (src/initMC.cc:126-196)
void initNuclearData(const Parameters& params, NuclearData& nd, MaterialDatabase& db);

} // namespace qsb
#endif
