// Parameters.hh -- input surface of the host model: the reference's command line + input deck.
//
// Same keys, defaults and precedence as the reference (src/Parameters.hh:99-177,
// src/Parameters.cc:80-95,230-255,384-418): the command line is read first, then the deck, and the
// deck wins; unknown keys are ignored; the last duplicate wins; Geometry blocks are kept in order.
// The implementation is a key table (one row per parameter) shared by the CLI reader, the deck
// reader and the echo, instead of the reference's three hand-written lists.
#ifndef QSB_PARAMETERS_HH
#define QSB_PARAMETERS_HH

#include <cstdint>
#include <map>
#include <string>
#include <vector>

namespace qsb {

struct GeometryParameters
{
    enum Shape { UNDEFINED, BRICK, SPHERE };
    std::string materialName;
    Shape  shape = UNDEFINED;
    double radius = 0, xCenter = 0, yCenter = 0, zCenter = 0;
    double xMin = 0, yMin = 0, zMin = 0, xMax = 0, yMax = 0, zMax = 0;
};

struct MaterialParameters
{
    std::string name;
    double mass = 1000.0;
    double totalCrossSection = 1.0;
    int    nIsotopes = 10;
    int    nReactions = 9;
    double sourceRate = 0.0;
    std::string scatteringCrossSection, absorptionCrossSection, fissionCrossSection;
    double scatteringCrossSectionRatio = 1.0, absorptionCrossSectionRatio = 1.0, fissionCrossSectionRatio = 1.0;
};

struct CrossSectionParameters
{
    std::string name;
    double aa = 0, bb = 0, cc = 0, dd = 0, ee = 1.0;
    double nuBar = 2.4;
};

struct SimulationParameters
{
    std::string inputFile, energySpectrum, crossSectionsOut;
    std::string boundaryCondition = "reflect";
    int      loadBalance = 0, cycleTimers = 0, debugThreads = 0;
    uint64_t nParticles = 1000000, batchSize = 0, nBatches = 10;
    int      nSteps = 10;
    int      nx = 10, ny = 10, nz = 10;
    int      seed = 1029384756;          // parsed and echoed, never used (as in the reference)
    int      xDom = 0, yDom = 0, zDom = 0;
    double   dt = 1e-8;
    double   fMax = 0.1;                 // parsed and echoed, never used
    double   lx = 100.0, ly = 100.0, lz = 100.0;
    double   eMin = 1e-9, eMax = 20;
    int      nGroups = 230;
    double   lowWeightCutoff = 0.001;
    int      balanceTallyReplications = 1, fluxTallyReplications = 1, cellTallyReplications = 1;
    int      coralBenchmark = 0;
};

struct Parameters
{
    SimulationParameters                          simulationParams;
    std::vector<GeometryParameters>               geometryParams;
    std::map<std::string, MaterialParameters>     materialParams;      // name order == material index order
    std::map<std::string, CrossSectionParameters> crossSectionParams;
};

// Throws std::runtime_error with a one-line reason on unusable input (missing deck, block without a
// name/shape); the C ABI turns that into QSB_ERR_INPUT.  argv[0] is the program name, as in main().
Parameters getParameters(int argc, const char* const* argv);
// Same as above but the deck is given as text instead of via -i <file>.
Parameters getParametersFromDeckText(int argc, const char* const* argv, const std::string& deckText);
std::string printParameters(const Parameters& params);
std::string commandLineHelp();

} // namespace qsb
#endif
