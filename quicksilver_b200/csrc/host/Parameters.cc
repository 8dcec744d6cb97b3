// Parameters.cc -- command line + deck reader (see Parameters.hh).
//
// Behaviour follows the reference: option table src/Parameters.cc:230-255, deck grammar
// src/parseUtils.cc:17-111 (blocks start at a column-0 "Name:" line, members are indented
// "key: value" or "key = value" lines, "//" starts a comment, blank lines are skipped, the first
// non-member line closes the block), value conversion by stream extraction (src/InputBlock.hh:38-49,
// so "100 // note" reads as 100 and a string value is its first word), default problem when no
// Geometry block is given (src/Parameters.cc:350-379).
#include "Parameters.hh"

#include <getopt.h>
#include <cctype>
#include <cstring>
#include <fstream>
#include <sstream>
#include <stdexcept>

namespace qsb {
namespace {

enum class Kind { Int, U64, Double, String, Flag };

struct SimKey
{
    const char* name;      // deck key and long option
    char        shortOpt;  // 0 = deck only
    Kind        kind;
    void*     (*field)(SimulationParameters&);
    const char* help;
};

#define QSB_FIELD(member) [](SimulationParameters& s) -> void* { return &s.member; }

// One row per simulation parameter.  shortOpt != 0 rows form the reference's 26-option command line.
const SimKey kSimKeys[] = {
    { "dt",               'D', Kind::Double, QSB_FIELD(dt),               "time step (seconds)" },
    { "fMax",             'f', Kind::Double, QSB_FIELD(fMax),             "max random mesh node displacement" },
    { "inputFile",        'i', Kind::String, QSB_FIELD(inputFile),        "name of input file" },
    { "energySpectrum",   'e', Kind::String, QSB_FIELD(energySpectrum),   "name of energy spectrum output file" },
    { "crossSectionsOut", 'S', Kind::String, QSB_FIELD(crossSectionsOut), "name of cross section output file" },
    { "loadBalance",      'l', Kind::Flag,   QSB_FIELD(loadBalance),      "enable/disable load balancing" },
    { "cycleTimers",      'c', Kind::Int,    QSB_FIELD(cycleTimers),      "enable/disable cycle timers" },
    { "debugThreads",     't', Kind::Int,    QSB_FIELD(debugThreads),     "set thread debug level to 1, 2, 3" },
    { "lx",               'X', Kind::Double, QSB_FIELD(lx),               "x-size of simulation (cm)" },
    { "ly",               'Y', Kind::Double, QSB_FIELD(ly),               "y-size of simulation (cm)" },
    { "lz",               'Z', Kind::Double, QSB_FIELD(lz),               "z-size of simulation (cm)" },
    { "nParticles",       'n', Kind::U64,    QSB_FIELD(nParticles),       "number of particles" },
    { "batchSize",        'g', Kind::U64,    QSB_FIELD(batchSize),        "number of particles in a vault/batch" },
    { "nBatches",         'b', Kind::U64,    QSB_FIELD(nBatches),         "number of vault/batch to start" },
    { "nSteps",           'N', Kind::Int,    QSB_FIELD(nSteps),           "number of time steps" },
    { "nx",               'x', Kind::Int,    QSB_FIELD(nx),               "number of mesh elements in x" },
    { "ny",               'y', Kind::Int,    QSB_FIELD(ny),               "number of mesh elements in y" },
    { "nz",               'z', Kind::Int,    QSB_FIELD(nz),               "number of mesh elements in z" },
    { "seed",             's', Kind::Int,    QSB_FIELD(seed),             "random number seed" },
    { "xDom",             'I', Kind::Int,    QSB_FIELD(xDom),             "number of domains (GPUs) in x" },
    { "yDom",             'J', Kind::Int,    QSB_FIELD(yDom),             "number of domains (GPUs) in y" },
    { "zDom",             'K', Kind::Int,    QSB_FIELD(zDom),             "number of domains (GPUs) in z" },
    { "bTally",           'B', Kind::Int,    QSB_FIELD(balanceTallyReplications), "number of balance tally replications" },
    { "fTally",           'F', Kind::Int,    QSB_FIELD(fluxTallyReplications),    "number of scalar flux tally replications" },
    { "cTally",           'C', Kind::Int,    QSB_FIELD(cellTallyReplications),    "number of scalar cell tally replications" },
    // deck-only keys (src/Parameters.cc:384-418)
    { "boundaryCondition", 0,  Kind::String, QSB_FIELD(boundaryCondition), "" },
    { "eMax",              0,  Kind::Double, QSB_FIELD(eMax),              "" },
    { "eMin",              0,  Kind::Double, QSB_FIELD(eMin),              "" },
    { "nGroups",           0,  Kind::Int,    QSB_FIELD(nGroups),           "" },
    { "lowWeightCutoff",   0,  Kind::Double, QSB_FIELD(lowWeightCutoff),   "" },
    { "coralBenchmark",    0,  Kind::Int,    QSB_FIELD(coralBenchmark),    "" },
};
#undef QSB_FIELD

// ---- deck -------------------------------------------------------------------------------------------

typedef std::map<std::string, std::string> KeyValues;
struct Block { std::string name; KeyValues kv; };

std::string trimmed(const std::string& s)
{
    size_t b = 0, e = s.size();
    while (b < e && isspace((unsigned char)s[b])) ++b;
    while (e > b && isspace((unsigned char)s[e - 1])) --e;
    return s.substr(b, e - b);
}

bool blankOrComment(std::string line)
{
    size_t c = line.find("//");
    if (c != std::string::npos) line.erase(c);
    return trimmed(line).empty();
}

// "  key: value" -> (indent, key, value); false when the line has no ':' or '=' after its indent.
bool splitLine(const std::string& line, int& indent, std::string& key, std::string& value)
{
    indent = 0;
    while (indent < (int)line.size() && isspace((unsigned char)line[indent])) ++indent;
    size_t delim = line.find_first_of(":=", indent);
    if (delim == std::string::npos) return false;
    key = trimmed(line.substr(indent, delim - indent));
    value = delim + 1 < line.size() ? trimmed(line.substr(delim + 1)) : std::string();
    return true;
}

std::vector<Block> readBlocks(std::istream& in)
{
    std::vector<std::string> lines;
    for (std::string line; std::getline(in, line);) lines.push_back(line);

    std::vector<Block> blocks;
    size_t i = 0;
    while (i < lines.size())
    {
        int indent; std::string key, value;
        if (!(splitLine(lines[i], indent, key, value) && indent == 0 && value.empty())) { ++i; continue; }
        blocks.push_back(Block{ key, {} });
        // members: until the first line that is neither blank/comment nor an indented key:value
        for (++i; i < lines.size(); ++i)
        {
            if (blankOrComment(lines[i])) continue;
            if (!splitLine(lines[i], indent, key, value) || indent == 0) break;   // re-examined as a block start
            blocks.back().kv[key] = value;
        }
    }
    return blocks;
}

template <typename T>
void extract(const KeyValues& kv, const char* key, T& out)
{
    KeyValues::const_iterator it = kv.find(key);
    if (it == kv.end() || it->second.empty()) return;        // "key:" with no value leaves the default
    std::istringstream s(it->second);
    T tmp = out;
    s >> tmp;
    if (!s) throw std::runtime_error(std::string("cannot parse value of '") + key + "': " + it->second);
    out = tmp;
}

void assign(const SimKey& k, SimulationParameters& sp, const KeyValues& kv)
{
    void* p = k.field(sp);
    switch (k.kind)
    {
        case Kind::Flag:
        case Kind::Int:    extract(kv, k.name, *static_cast<int*>(p)); break;
        case Kind::U64:    extract(kv, k.name, *static_cast<uint64_t*>(p)); break;
        case Kind::Double: extract(kv, k.name, *static_cast<double*>(p)); break;
        case Kind::String: extract(kv, k.name, *static_cast<std::string*>(p)); break;
    }
}

void scanSimulation(const KeyValues& kv, Parameters& pp)
{
    for (const SimKey& k : kSimKeys)
        if (std::strcmp(k.name, "inputFile") != 0)      // the deck cannot redirect to another deck
            assign(k, pp.simulationParams, kv);
}

void scanGeometry(const KeyValues& kv, Parameters& pp)
{
    GeometryParameters g;
    extract(kv, "material", g.materialName);
    std::string shape;
    extract(kv, "shape", shape);
    if (shape == "brick")
    {
        g.shape = GeometryParameters::BRICK;
        extract(kv, "xMax", g.xMax); extract(kv, "xMin", g.xMin);
        extract(kv, "yMax", g.yMax); extract(kv, "yMin", g.yMin);
        extract(kv, "zMax", g.zMax); extract(kv, "zMin", g.zMin);
    }
    else if (shape == "sphere")
    {
        g.shape = GeometryParameters::SPHERE;
        extract(kv, "radius", g.radius);
        extract(kv, "xCenter", g.xCenter); extract(kv, "yCenter", g.yCenter); extract(kv, "zCenter", g.zCenter);
    }
    else
        throw std::runtime_error("Geometry block needs shape: brick | sphere");
    pp.geometryParams.push_back(g);
}

void scanMaterial(const KeyValues& kv, Parameters& pp)
{
    std::string name;
    extract(kv, "name", name);
    if (name.empty()) throw std::runtime_error("Material block without a name");
    MaterialParameters& m = pp.materialParams[name];
    m.name = name;
    extract(kv, "mass", m.mass);
    extract(kv, "nIsotopes", m.nIsotopes);
    extract(kv, "nReactions", m.nReactions);
    extract(kv, "sourceRate", m.sourceRate);
    extract(kv, "totalCrossSection", m.totalCrossSection);
    extract(kv, "absorptionCrossSection", m.absorptionCrossSection);
    extract(kv, "fissionCrossSection", m.fissionCrossSection);
    extract(kv, "scatteringCrossSection", m.scatteringCrossSection);
    extract(kv, "absorptionCrossSectionRatio", m.absorptionCrossSectionRatio);
    extract(kv, "fissionCrossSectionRatio", m.fissionCrossSectionRatio);
    extract(kv, "scatteringCrossSectionRatio", m.scatteringCrossSectionRatio);
}

void scanCrossSection(const KeyValues& kv, Parameters& pp)
{
    std::string name;
    extract(kv, "name", name);
    if (name.empty()) throw std::runtime_error("CrossSection block without a name");
    CrossSectionParameters& c = pp.crossSectionParams[name];
    c.name = name;
    extract(kv, "A", c.aa); extract(kv, "B", c.bb); extract(kv, "C", c.cc);
    extract(kv, "D", c.dd); extract(kv, "E", c.ee);
    extract(kv, "nuBar", c.nuBar);
}

void applyDeck(std::istream& in, Parameters& pp)
{
    for (const Block& b : readBlocks(in))
    {
        if      (b.name == "Simulation")   scanSimulation(b.kv, pp);
        else if (b.name == "Geometry")     scanGeometry(b.kv, pp);
        else if (b.name == "Material")     scanMaterial(b.kv, pp);
        else if (b.name == "CrossSection") scanCrossSection(b.kv, pp);
    }
}

// ---- command line -----------------------------------------------------------------------------------

void parseCommandLine(int argc, const char* const* argv, Parameters& pp, bool& help)
{
    SimulationParameters& sp = pp.simulationParams;
    std::vector<option> longOpts;
    std::string shortOpts;
    for (const SimKey& k : kSimKeys)
    {
        if (!k.shortOpt) continue;
        const int hasArg = k.kind == Kind::Flag ? no_argument : required_argument;
        longOpts.push_back(option{ k.name, hasArg, nullptr, k.shortOpt });
        shortOpts += k.shortOpt;
        if (hasArg) shortOpts += ':';
    }
    longOpts.push_back(option{ "help", no_argument, nullptr, 'h' });
    shortOpts += 'h';
    longOpts.push_back(option{ nullptr, 0, nullptr, 0 });

    std::vector<std::string> store(argv, argv + argc);
    std::vector<char*> args;
    for (std::string& s : store) args.push_back(&s[0]);
    args.push_back(nullptr);

    optind = 0;   // glibc: full re-initialisation, the library may parse many command lines per process
    opterr = 0;
    for (int c; (c = getopt_long(argc, args.data(), shortOpts.c_str(), longOpts.data(), nullptr)) != -1;)
    {
        if (c == 'h') { help = true; continue; }
        const SimKey* key = nullptr;
        for (const SimKey& k : kSimKeys) if (k.shortOpt == c) key = &k;
        if (!key) continue;                                   // unknown switch: ignored, like the reference
        void* p = key->field(sp);
        switch (key->kind)                                    // sscanf semantics: no type checking
        {
            case Kind::Flag:   *static_cast<int*>(p) = 1; break;
            case Kind::Int:    sscanf(optarg, "%d", static_cast<int*>(p)); break;
            case Kind::U64:    { unsigned long long v; if (sscanf(optarg, "%llu", &v) == 1) *static_cast<uint64_t*>(p) = v; } break;
            case Kind::Double: sscanf(optarg, "%lf", static_cast<double*>(p)); break;
            case Kind::String: *static_cast<std::string*>(p) = optarg; break;
        }
    }
}

// If no geometry was given the user gets the built-in one-material problem (src/Parameters.cc:350-379).
void supplyDefaults(Parameters& pp)
{
    if (!pp.geometryParams.empty()) return;
    CrossSectionParameters flat;
    flat.name = "flat";
    pp.crossSectionParams[flat.name] = flat;

    MaterialParameters m;
    m.name = "sourceMaterial";
    m.mass = 1000.0;
    m.sourceRate = 1e10;
    m.scatteringCrossSection = m.absorptionCrossSection = m.fissionCrossSection = "flat";
    m.fissionCrossSectionRatio = 0.1;
    pp.materialParams[m.name] = m;

    GeometryParameters g;
    g.materialName = "sourceMaterial";
    g.shape = GeometryParameters::BRICK;
    g.xMax = pp.simulationParams.lx;
    g.yMax = pp.simulationParams.ly;
    g.zMax = pp.simulationParams.lz;
    pp.geometryParams.push_back(g);
}

Parameters build(int argc, const char* const* argv, const std::string* deckText)
{
    Parameters pp;
    bool help = false;
    parseCommandLine(argc, argv, pp, help);
    if (help) throw std::runtime_error(commandLineHelp());

    // a spectrum / cross-section file named on the command line survives the deck (src/Parameters.cc:85-91)
    const std::string cliSpectrum = pp.simulationParams.energySpectrum;
    const std::string cliXsOut = pp.simulationParams.crossSectionsOut;
    if (deckText)
    {
        std::istringstream in(*deckText);
        applyDeck(in, pp);
    }
    else if (!pp.simulationParams.inputFile.empty())
    {
        std::ifstream in(pp.simulationParams.inputFile.c_str());
        if (!in) throw std::runtime_error("cannot open input file " + pp.simulationParams.inputFile);
        applyDeck(in, pp);
    }
    if (!cliSpectrum.empty()) pp.simulationParams.energySpectrum = cliSpectrum;
    if (!cliXsOut.empty())    pp.simulationParams.crossSectionsOut = cliXsOut;
    supplyDefaults(pp);
    return pp;
}

} // namespace

Parameters getParameters(int argc, const char* const* argv) { return build(argc, argv, nullptr); }

Parameters getParametersFromDeckText(int argc, const char* const* argv, const std::string& deckText)
{
    return build(argc, argv, &deckText);
}

std::string commandLineHelp()
{
    std::ostringstream out;
    out << "\n  Arguments are: \n";
    for (const SimKey& k : kSimKeys)
        if (k.shortOpt)
            out << "   --" << k.name << std::string(18 - std::min<size_t>(18, strlen(k.name)), ' ') << " -" << k.shortOpt
                << "  arg=" << (k.kind == Kind::Flag ? 0 : 1) << "  " << k.help << "\n";
    out << "\n";
    return out.str();
}

// Echo in deck syntax, key order of the reference's printParameters (src/Parameters.cc:97-215).
std::string printParameters(const Parameters& pp)
{
    const SimulationParameters& s = pp.simulationParams;
    std::ostringstream out;
    out << "Simulation:\n"
        << "   dt: " << s.dt << "\n"
        << "   fMax: " << s.fMax << "\n"
        << "   inputFile: " << s.inputFile << "\n"
        << "   energySpectrum: " << s.energySpectrum << "\n"
        << "   boundaryCondition: " << s.boundaryCondition << "\n"
        << "   loadBalance: " << s.loadBalance << "\n"
        << "   cycleTimers: " << s.cycleTimers << "\n"
        << "   debugThreads: " << s.debugThreads << "\n"
        << "   lx: " << s.lx << "\n" << "   ly: " << s.ly << "\n" << "   lz: " << s.lz << "\n"
        << "   nParticles: " << s.nParticles << "\n"
        << "   batchSize: " << s.batchSize << "\n"
        << "   nBatches: " << s.nBatches << "\n"
        << "   nSteps: " << s.nSteps << "\n"
        << "   nx: " << s.nx << "\n" << "   ny: " << s.ny << "\n" << "   nz: " << s.nz << "\n"
        << "   seed: " << s.seed << "\n"
        << "   xDom: " << s.xDom << "\n" << "   yDom: " << s.yDom << "\n" << "   zDom: " << s.zDom << "\n"
        << "   eMax: " << s.eMax << "\n" << "   eMin: " << s.eMin << "\n"
        << "   nGroups: " << s.nGroups << "\n"
        << "   lowWeightCutoff: " << s.lowWeightCutoff << "\n"
        << "   bTally: " << s.balanceTallyReplications << "\n"
        << "   fTally: " << s.fluxTallyReplications << "\n"
        << "   cTally: " << s.cellTallyReplications << "\n"
        << "   coralBenchmark: " << s.coralBenchmark << "\n"
        << "   crossSectionsOut:" << s.crossSectionsOut << "\n\n";
    for (const GeometryParameters& g : pp.geometryParams)
    {
        out << "Geometry:\n   material: " << g.materialName << "\n";
        if (g.shape == GeometryParameters::BRICK)
            out << "   shape: brick\n"
                << "   xMax: " << g.xMax << "\n   xMin: " << g.xMin << "\n"
                << "   yMax: " << g.yMax << "\n   yMin: " << g.yMin << "\n"
                << "   zMax: " << g.zMax << "\n   zMin: " << g.zMin << "\n";
        else
            out << "   shape: sphere\n"
                << "   xCenter: " << g.xCenter << "\n   yCenter: " << g.yCenter << "\n   zCenter: " << g.zCenter << "\n";
        out << "\n";
    }
    for (const auto& kv : pp.materialParams)
    {
        const MaterialParameters& m = kv.second;
        out << "Material:\n"
            << "   name: " << m.name << "\n"
            << "   mass: " << m.mass << "\n"
            << "   nIsotopes: " << m.nIsotopes << "\n"
            << "   nReactions: " << m.nReactions << "\n"
            << "   sourceRate: " << m.sourceRate << "\n"
            << "   totalCrossSection: " << m.totalCrossSection << "\n"
            << "   absorptionCrossSection: " << m.absorptionCrossSection << "\n"
            << "   fissionCrossSection: " << m.fissionCrossSection << "\n"
            << "   scatteringCrossSection: " << m.scatteringCrossSection << "\n"
            << "   absorptionCrossSectionRatio: " << m.absorptionCrossSectionRatio << "\n"
            << "   fissionCrossSectionRatio: " << m.fissionCrossSectionRatio << "\n"
            << "   scatteringCrossSectionRatio: " << m.scatteringCrossSectionRatio << "\n\n";
    }
    for (const auto& kv : pp.crossSectionParams)
    {
        const CrossSectionParameters& c = kv.second;
        out << "CrossSection:\n"
            << "   name: " << c.name << "\n"
            << "   A: " << c.aa << "\n   B: " << c.bb << "\n   C: " << c.cc << "\n   D: " << c.dd << "\n   E: " << c.ee << "\n"
            << "   nuBar: " << c.nuBar << "\n";
    }
    return out.str();
}

} // namespace qsb
