// NuclearData.cc -- see NuclearData.hh.  All transcendental calls (log, exp, log10, pow) are host
// libm, the same library the reference uses, so the tables match the reference bit for bit.
#include "NuclearData.hh"

#include <cmath>
#include <stdexcept>

namespace qsb {

// Edges are log-spaced with (numGroups + 1) in the denominator -- the reference's spacing, kept as is
// (src/NuclearData.cc:105-119) -- and the last edge is pinned to energyHigh.
NuclearData::NuclearData(int nGroups, double energyLow, double energyHigh)
: numGroups(nGroups), energies(nGroups + 1)
{
    if (!(energyLow < energyHigh) || nGroups < 1) throw std::runtime_error("NuclearData: need eMin < eMax and nGroups >= 1");
    energies[0] = energyLow;
    energies[nGroups] = energyHigh;
    const double logLow = std::log(energyLow);
    const double logHigh = std::log(energyHigh);
    const double delta = (logHigh - logLow) / (nGroups + 1.0);
    for (int i = 1; i < nGroups; ++i)
        energies[i] = std::exp(logLow + delta * i);
}

namespace {
double poly(const CrossSectionParameters& c, double x)
{
    return c.ee + x * (c.dd + x * (c.cc + x * (c.bb + x * (c.aa))));
}
}

int NuclearData::addIsotope(int nReactions, const CrossSectionParameters& fission, const CrossSectionParameters& scatter,
                            const CrossSectionParameters& absorption, double nuBar, double totalXs,
                            double fissionWeight, double scatterWeight, double absorptionWeight)
{
    Isotope iso;
    iso.nReactions = nReactions;
    iso.sigmaOffset = sigma.size();
    iso.nuBar = nuBar;
    iso.reactionType.resize(nReactions);
    sigma.resize(sigma.size() + (size_t)nReactions * numGroups);

    const double totalWeight = fissionWeight + scatterWeight + absorptionWeight;
    int nFission = nReactions / 3, nScatter = nReactions / 3, nAbsorption = nReactions / 3;
    if (nReactions % 3 >= 1) ++nScatter;
    if (nReactions % 3 == 2) ++nFission;
    const double fissionXs    = (totalXs * fissionWeight)    / (nFission    * totalWeight);
    const double scatterXs    = (totalXs * scatterWeight)    / (nScatter    * totalWeight);
    const double absorptionXs = (totalXs * absorptionWeight) / (nAbsorption * totalWeight);

    for (int r = 0; r < nReactions; ++r)
    {
        const CrossSectionParameters* shape; double target;
        switch (r % 3)
        {
            case 0:  iso.reactionType[r] = ReactionType::Scatter;    shape = &scatter;    target = scatterXs;    break;
            case 1:  iso.reactionType[r] = ReactionType::Fission;    shape = &fission;    target = fissionXs;    break;
            default: iso.reactionType[r] = ReactionType::Absorption; shape = &absorption; target = absorptionXs; break;
        }
        double* xs = &sigma[iso.sigmaOffset + (size_t)r * numGroups];
        for (int g = 0; g < numGroups; ++g)
        {
            const double mid = (energies[g] + energies[g + 1]) / 2.0;
            xs[g] = std::pow(10, poly(*shape, std::log10(mid)));
        }
        double normalization = 0.0;                      // value of the group that contains 1 MeV
        for (int g = 0; g < numGroups; ++g)
            if (energies[g + 1] >= 1.0) { normalization = xs[g]; break; }
        if (!(normalization > 0.0)) throw std::runtime_error("NuclearData: no energy group contains 1 MeV");
        const double scale = target / normalization;
        for (int g = 0; g < numGroups; ++g) xs[g] *= scale;
    }
    isotopes.push_back(iso);
    return (int)isotopes.size() - 1;
}

int NuclearData::getEnergyGroup(double energy) const
{
    const int n = (int)energies.size();
    if (energy <= energies[0]) return 0;
    if (energy > energies[n - 1]) return n - 1;
    int lo = 0, hi = n - 1;
    while (hi != lo + 1)
    {
        const int mid = (hi + lo) / 2;
        if (energy < energies[mid]) hi = mid; else lo = mid;
    }
    return lo;
}

double NuclearData::totalCrossSection(int iso, int group) const
{
    double total = 0.0;
    for (int r = 0; r < isotopes[iso].nReactions; ++r) total += sigmaOf(iso, r, group);
    return total;
}

int MaterialDatabase::findMaterial(const std::string& name) const
{
    for (size_t i = 0; i < mat.size(); ++i) if (mat[i].name == name) return (int)i;
    return -1;
}

void initNuclearData(const Parameters& params, NuclearData& nd, MaterialDatabase& db)
{
    for (const auto& kv : params.materialParams)           // std::map: alphabetical == material index order
    {
        const MaterialParameters& mp = kv.second;
        auto xs = [&](const std::string& n) -> const CrossSectionParameters& {
            auto it = params.crossSectionParams.find(n);
            if (it == params.crossSectionParams.end())
                throw std::runtime_error("material " + mp.name + " names unknown cross section '" + n + "'");
            return it->second;
        };
        Material m;
        m.name = mp.name;
        m.mass = mp.mass;
        m.sourceRate = mp.sourceRate;
        m.nuBar = xs(mp.fissionCrossSection).nuBar;
        for (int i = 0; i < mp.nIsotopes; ++i)
        {
            const int gid = nd.addIsotope(mp.nReactions, xs(mp.fissionCrossSection), xs(mp.scatteringCrossSection),
                                          xs(mp.absorptionCrossSection), m.nuBar, mp.totalCrossSection,
                                          mp.fissionCrossSectionRatio, mp.scatteringCrossSectionRatio,
                                          mp.absorptionCrossSectionRatio);
            m.isoGid.push_back(gid);
            m.atomFraction.push_back(1.0 / mp.nIsotopes);
        }
        db.mat.push_back(m);
    }
}

} // namespace qsb
