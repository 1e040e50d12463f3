// tools.cpp -- the reference's hot-path commands, rewritten over the batched host classes of theseed.hpp.
//
//   kmers_reps repdb      RepGenomeDbProcessor.java      create / load a rep DB, report closest reps of stdin FASTA
//   kmers_reps distances  DistanceMatrixProcessor.java   all-vs-all distances of the reps of a DB
//   kmers_reps group      RepMatrixProcessor.java        incremental all-vs-all + greedy grouping by distance
//   kmers_reps classify   ClassifyProcessor.java         min distance of each input sequence to its named group
//   kmers_reps gtoclass   GtoClassProcessor.java          closest representative of every genome of a GTO directory (BASELINE configs[0])
//   kmers_reps gtoreps    GtoRepGenomeProcessor.java      closest representative of every genome; --create / --update add the outliers
//   kmers_reps buildfromgtos GtoRepBuildProcessor.java     greedy representative set from a GTO directory, optional --verify report
//   kmers_reps neighbors  NeighborProcessor.java          the N closest representatives of each input genome, ties retained
//   kmers_reps kmerops    TargetProcessor.java:160-195 (the SequenceKmers set members: retainAll / removeAll / iteration)
//   kmers_reps buildsets  GenomeProcessor.java + BaseGenomeProcessor.java:151-214,348-407,433-442
//                                                       the hot part of the `genomes` pipeline: hierarchical rep sets at
//                                                       several thresholds, saved as rep<thr>.ser + rep<thr>.list.tbl
//
// Same options, same output text.  What changed is only what north_star lists: the
// per-pair / per-query loops became batched calls; every similarity comes from the GPU library.
#include "genome.hpp"

#include <chrono>
#include <cmath>

using namespace theseed;

static void usage() {
    std::cerr << "usage: kmers_reps <repdb|distances|group|classify> [options]\n"
                 "  repdb     [-m sim] [-K k] [-p role] [--create fasta] [-x maxErr] [--list file] [--stats file] [--null] [--batch n] dbFile < input.fa\n"
                 "  distances dbFile\n"
                 "  group     [-K k] [-i fasta] [-m maxDist]\n"
                 "  classify  [-K k] [-i fasta] [-m maxDist] controlFile\n"
                 "  gtoclass  genomeDir repgen1.db [repgen2.db ...]\n"
                 "  gtoreps   [--create thr] [--update newDb] [--filter ids.tbl] [--blackList ids.tbl] [--rna] [-o out] repDb gtoDir\n"
                 "  buildfromgtos [-K k] [-m sim] [--verify seeds.fa] [--batch n] gtoDir roleFile outDb\n"
                 "  neighbors [-n num] [--min sim] dbFile seeds.tbl\n"
                 "  buildsets [-K k] [-t 10,50,100,200] [-p role] [--batch n] outDir input.fa\n";
}

struct Args {
    std::map<std::string, std::string> opt;
    std::set<std::string> flag;
    std::vector<std::string> pos;
};
// aliases -> canonical option name; value-taking options listed in `valued`
static Args parse(int argc, char **argv, int first, const std::map<std::string, std::string> &alias, const std::set<std::string> &valued) {
    Args a;
    for (int i = first; i < argc; ++i) {
        std::string s = argv[i];
        if (s.size() > 1 && s[0] == '-' && !(s.size() > 1 && std::isdigit((unsigned char)s[1]))) {
            auto it = alias.find(s);
            if (it == alias.end()) throw std::runtime_error("ParseFailureException: \"" + s + "\" is not a valid option");
            if (valued.count(it->second)) {
                if (i + 1 >= argc) throw std::runtime_error("ParseFailureException: option " + s + " takes an operand");
                a.opt[it->second] = argv[++i];
            } else a.flag.insert(it->second);
        } else a.pos.push_back(s);
    }
    return a;
}

static double now_s() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

// ---- repdb (RepGenomeDbProcessor.java:119-283) ------------------------------------------------------
static int cmd_repdb(int argc, char **argv) {
    Args a = parse(argc, argv, 2,
                   {{"-m", "m"}, {"--sim", "m"}, {"--minScore", "m"}, {"-K", "K"}, {"--kmer", "K"}, {"-p", "p"}, {"--prot", "p"}, {"--role", "p"},
                    {"--create", "create"}, {"--list", "list"}, {"--stats", "stats"}, {"--null", "null"}, {"-x", "x"}, {"--maxErr", "x"},
                    {"--dna", "dna"}, {"--batch", "batch"}, {"-v", "v"}, {"--verbose", "v"}},
                   {"m", "K", "p", "create", "list", "stats", "x", "batch"});
    if (a.pos.size() != 1) throw std::runtime_error("ParseFailureException: Argument \"repDbFile\" is required");
    if (a.flag.count("dna")) throw std::runtime_error("--dna is outside the accelerated path (DnaKmers, SURVEY.md 8f-4)");
    const std::string dbFile = a.pos[0];
    int threshold = a.opt.count("m") ? std::stoi(a.opt["m"]) : 100;      // setDefaults :183
    if (a.opt.count("K")) ProteinKmers::setKmerSize(std::stoi(a.opt["K"]));
    std::string protName = a.opt.count("p") ? a.opt["p"] : RepGenomeDb::DEFAULT_PROTEIN;
    int maxErr = a.opt.count("x") ? std::stoi(a.opt["x"]) : 20;
    uint32_t batch = a.opt.count("batch") ? (uint32_t)std::stoul(a.opt["batch"]) : 0;
    const bool verbose = a.flag.count("v") != 0;
    std::unique_ptr<RepGenomeDb> repDB;
    double t0 = now_s();
    if (a.opt.count("create")) {
        // createRepDb :119-153: candidates in file order, sequences with >= maxErr 'X' deferred to a second pass
        repDB = std::make_unique<RepGenomeDb>(threshold, protName);
        FastaInputStream in(a.opt["create"]);
        std::vector<RepGenome> order, deferred;
        Sequence s;
        while (in.next(s)) {
            long bad = std::count(s.sequence.begin(), s.sequence.end(), 'X');
            (bad >= maxErr ? deferred : order).emplace_back(s);
        }
        size_t nmain = order.size();
        for (auto &d : deferred) order.push_back(std::move(d));
        auto added = repDB->checkGenomes(order, batch);
        if (verbose)
            std::cerr << order.size() << " total genomes input (" << order.size() - nmain << " deferred), " << repDB->size() << " kept ("
                      << (long)(order.size() / std::max(1e-9, now_s() - t0)) << " genomes/second).\n";
        repDB->save(dbFile);
    } else {
        std::ifstream probe(dbFile);
        if (!probe) throw std::runtime_error("java.io.FileNotFoundException: Create not specified and specified repDB file not found.");
        repDB = RepGenomeDb::load(dbFile);
    }
    if (a.opt.count("list")) {   // :212-222
        std::ofstream ls(a.opt["list"]);
        ls << "genome_id\tname\n";
        for (const RepGenome *r : repDB->sorted()) ls << r->getGenomeId() << '\t' << r->getName() << '\n';
    }
    if (!a.flag.count("null")) {   // :224-262
        std::cout << "genome_id\tname\trep_id\tscore\n";
        FastaInputStream in(std::cin);
        std::vector<Sequence> seqs = in.readAll();
        std::vector<ProteinKmers> qs;
        for (auto &q : seqs) qs.emplace_back(q.sequence);
        std::vector<const ProteinKmers *> qp;
        for (auto &q : qs) qp.push_back(&q);
        t0 = now_s();
        auto res = repDB->findClosest(qp.begin(), qp.end());
        std::map<const RepGenome *, std::pair<int, int>> counts;   // QualityCountMap: good, bad
        int outliers = 0;
        for (size_t i = 0; i < seqs.size(); ++i) {
            const auto &r = res[i];
            if (!r.isRepresented()) { ++outliers; if (r.getRepresentative()) counts[r.getRepresentative()].second++; }
            else counts[r.getRepresentative()].first++;
            int sim = r.getSimilarity();
            std::cout << genomeOf(seqs[i].label) << '\t' << seqs[i].comment << '\t' << r.getGenomeId() << '\t'
                      << (sim == SequenceKmers::INFINITY_ ? std::string("MAX") : std::to_string(sim)) << '\n';
        }
        if (verbose)
            std::cerr << seqs.size() << " genomes processed, " << outliers << " outliers, "
                      << (long)(seqs.size() / std::max(1e-9, now_s() - t0)) << " genomes/second.\n";
        if (a.opt.count("stats")) {   // :263-278; bestKeys(): most represented first, ties by genome id
            std::ofstream st(a.opt["stats"]);
            st << "rep_id\trep_name\trepresented\toutliers\n";
            std::vector<std::pair<const RepGenome *, std::pair<int, int>>> v(counts.begin(), counts.end());
            std::sort(v.begin(), v.end(), [](auto &x, auto &y) {
                if (x.second.first != y.second.first) return x.second.first > y.second.first;
                if (x.second.second != y.second.second) return x.second.second < y.second.second;
                return x.first->compareTo(*y.first) < 0;
            });
            for (auto &e : v) st << e.first->getGenomeId() << '\t' << e.first->getName() << '\t' << e.second.first << '\t' << e.second.second << '\n';
        }
    }
    return 0;
}

// ---- distances (DistanceMatrixProcessor.java:54-101) -----------------------------------------------
static int cmd_distances(int argc, char **argv) {
    Args a = parse(argc, argv, 2, {{"-v", "v"}, {"--verbose", "v"}}, {});
    if (a.pos.size() != 1) throw std::runtime_error("ParseFailureException: Argument \"repDbFile\" is required");
    auto repDB = RepGenomeDb::load(a.pos[0]);
    std::cout << "genome_1\tgenome_2\tdistance\tsim\n";
    auto reps = repDB->all();
    double t0 = now_s();
    SimilarityMatrix m = SimilarityMatrix::compute(reps.begin(), reps.end());
    long processed = 0, zeroes = 0;
    double distSum = 0, distSqSum = 0, simSum = 0, simSqSum = 0;
    std::string line;
    for (size_t i = 0; i < reps.size(); ++i)
        for (size_t j = 0; j < reps.size(); ++j)
            if (reps[i]->compareTo(*reps[j]) < 0) {   // each unordered pair once, in the reference's loop order
                double distance = m.distance(i, j);
                int sim = m.similarity(i, j);
                distSum += distance; distSqSum += distance * distance; simSum += sim; simSqSum += (double)sim * sim;
                if (sim == 0) zeroes++;
                std::cout << reps[i]->getGenomeId() << '\t' << reps[j]->getGenomeId() << '\t' << java::format_fixed(distance, 8, 6) << '\t' << sim << '\n';
                processed++;
            }
    if (a.flag.count("v")) {
        double distMean = distSum / processed, simMean = simSum / processed;
        std::cerr << processed << " pairs processed in " << (now_s() - t0) << " seconds.\n"
                  << "Mean distance is " << java::double_to_string(distMean) << " with standard deviation "
                  << java::double_to_string(std::sqrt(distSqSum / processed - distMean * distMean)) << ".\n"
                  << "Mean similarity is " << java::double_to_string(simMean) << " with standard deviation "
                  << java::double_to_string(std::sqrt(simSqSum / processed - simMean * simMean)) << ".\n"
                  << zeroes << " pairs had 0 similarity.\n";
    }
    return 0;
}

// ---- group (RepMatrixProcessor.java:83-130 + SequenceInfo.java) -----------------------------------------
namespace {
struct SequenceInfo {   // SequenceInfo.java:30-149: running distance statistics in the reference's order
    Sequence original;
    int distCount = 0;
    double distSum = 0, distSquareSum = 0, distMax = 0, distMin = 1.7976931348623157e308;
    void addDistance(double dist) {   // :83-89
        distCount++; distSum += dist; distSquareSum += dist * dist;
        if (dist > distMax) distMax = dist;
        if (dist < distMin) distMin = dist;
    }
    double getSDev() const {   // :103-110
        double Nm1 = distCount - 1, N = distCount;
        return N > 1.0 ? std::sqrt(distSquareSum / Nm1 - distSum * distSum / (N * Nm1)) : 0.0;
    }
};
int double_compare(double a, double b) { return a < b ? -1 : a > b ? 1 : 0; }   // Double.compare on finite values
}  // namespace

static int cmd_group(int argc, char **argv) {
    Args a = parse(argc, argv, 2, {{"-K", "K"}, {"--kmer", "K"}, {"-i", "i"}, {"--input", "i"}, {"-m", "m"}, {"--max", "m"}, {"--maxDist", "m"}, {"-v", "v"}},
                   {"K", "i", "m"});
    ProteinKmers::setKmerSize(a.opt.count("K") ? std::stoi(a.opt["K"]) : 8);
    double maxDist = a.opt.count("m") ? std::stod(a.opt["m"]) : 0.995;
    std::vector<Sequence> seqs = a.opt.count("i") ? FastaInputStream(a.opt["i"]).readAll() : FastaInputStream(std::cin).readAll();
    std::vector<ProteinKmers> km;
    for (auto &s : seqs) km.emplace_back(s.sequence);
    SimilarityMatrix m = SimilarityMatrix::compute(km.begin(), km.end());
    std::vector<SequenceInfo> info(seqs.size());
    for (size_t i = 0; i < seqs.size(); ++i) {   // :88-95: each new sequence against all previous ones, in order
        info[i].original = seqs[i];
        for (size_t j = 0; j < i; ++j) { double d = m.distance(i, j); info[i].addDistance(d); info[j].addDistance(d); }
    }
    std::vector<size_t> sorted(seqs.size());
    for (size_t i = 0; i < sorted.size(); ++i) sorted[i] = i;
    std::stable_sort(sorted.begin(), sorted.end(), [&](size_t x, size_t y) {   // SequenceInfo.compareTo :140-148
        int c = double_compare(info[x].distSum, info[y].distSum);
        if (c == 0) c = double_compare(info[x].getSDev(), info[y].getSDev());
        if (c == 0) c = info[x].original.label.compare(info[y].original.label);
        return c < 0;
    });
    FastaOutputStream out(std::cout);
    std::vector<size_t> kept;
    for (size_t x : sorted) {   // :102-124
        if (info[x].distMin > maxDist) continue;   // singleton group: not written
        bool found = false;
        for (size_t y : kept) if (m.distance(x, y) <= maxDist) { found = true; break; }
        if (!found) { kept.push_back(x); out.write(info[x].original); }
    }
    std::cerr << kept.size() << " sequences form the representation set.\n";
    return 0;
}

// ---- classify (ClassifyProcessor.java:99-126) ---------------------------------------------------------
static int cmd_classify(int argc, char **argv) {
    Args a = parse(argc, argv, 2, {{"-K", "K"}, {"--kmer", "K"}, {"-i", "i"}, {"--input", "i"}, {"-m", "m"}, {"--max", "m"}, {"--maxDist", "m"}, {"-v", "v"}},
                   {"K", "i", "m"});
    if (a.pos.size() != 1) throw std::runtime_error("ParseFailureException: Argument \"controlFile\" is required");
    ProteinKmers::setKmerSize(a.opt.count("K") ? std::stoi(a.opt["K"]) : 8);
    double maxDist = a.opt.count("m") ? std::stod(a.opt["m"]) : 0.995;
    KmerCollectionGroup groups;
    std::ifstream ctl(a.pos[0]);
    if (!ctl) throw std::runtime_error("java.io.FileNotFoundException: " + a.pos[0]);
    std::string line;
    while (std::getline(ctl, line)) {   // headerless, two tab-separated fields: group, FASTA file
        if (!line.empty() && line.back() == '\r') line.pop_back();
        if (line.empty()) continue;
        size_t tab = line.find('\t');
        if (tab == std::string::npos) throw std::runtime_error("java.io.IOException: control file line needs 2 fields: " + line);
        FastaInputStream fs(line.substr(tab + 1));
        Sequence s;
        while (fs.next(s)) groups.addSequence(s, line.substr(0, tab));
    }
    std::cout << "seq_id\tgroup\tdistance\n";
    std::vector<Sequence> seqs = a.opt.count("i") ? FastaInputStream(a.opt["i"]).readAll() : FastaInputStream(std::cin).readAll();
    std::vector<std::pair<const Sequence *, std::string>> q;
    for (auto &s : seqs) q.emplace_back(&s, s.comment);
    std::vector<double> d = groups.getDistances(q);
    for (size_t i = 0; i < seqs.size(); ++i)
        if (d[i] <= maxDist) std::cout << seqs[i].label << '\t' << seqs[i].comment << '\t' << java::format_fixed(d[i], 12, 8) << '\n';
    return 0;
}

// ---- gtoclass (GtoClassProcessor.java:100-186) ----------------------------------------------------------------
// For every genome of the source directory (Bacteria and Archaea only, :118) the seed protein is looked up and its closest
// representative in each RepGen database is reported, blank when it is not represented (:133-141).  The per-genome
// findClosest calls become one batched call per database.
static int cmd_gtoclass(int argc, char **argv) {
    Args a = parse(argc, argv, 2, {{"-v", "v"}, {"--verbose", "v"}, {"-t", "t"}, {"--type", "t"}}, {"t"});
    if (a.pos.size() < 2) throw std::runtime_error("ParseFailureException: Arguments \"genomeDir\" and \"repgen files\" are required");
    if (a.opt.count("t") && a.opt["t"] != "DIR") throw std::runtime_error("ParseFailureException: only genome sources of type DIR are read here");
    std::vector<std::unique_ptr<RepGenomeDb>> repGens;
    std::string header = "genome_id\tgenome_name";
    for (size_t i = 1; i < a.pos.size(); ++i) {   // buildRepGens :166-186
        std::ifstream probe(a.pos[i]);
        if (!probe) throw std::runtime_error("java.io.FileNotFoundException: RepGen file " + a.pos[i] + " is not found or unreadable.");
        repGens.push_back(RepGenomeDb::load(a.pos[i]));
        header += "\tRepGen." + std::to_string(repGens.back()->getThreshold());
    }
    std::cout << header << '\n';
    struct Row { std::string id, name; };
    std::vector<Row> rows; std::vector<ProteinKmers> seeds;
    int gCount = 0, noSeed = 0;
    for (const std::string &file : genomeDirectory(a.pos[0])) {
        Genome genome(file);
        ++gCount;
        if (genome.getDomain() != "Bacteria" && genome.getDomain() != "Archaea") continue;   // LEGAL_DOMAINS :63,118
        const std::string protId = findSeed(genome);
        if (protId.empty()) { ++noSeed; continue; }                                           // :123-125
        rows.push_back({genome.getId(), genome.getName()});
        seeds.emplace_back(genome.getFeature(protId)->getProteinTranslation());
    }
    std::vector<const ProteinKmers *> qp;
    for (auto &q : seeds) qp.push_back(&q);
    std::vector<std::vector<RepGenomeDb::Representation>> res;
    for (auto &db : repGens) res.push_back(db->findClosest(qp.begin(), qp.end()));
    int foundCount = 0, errorCount = noSeed * (int)repGens.size();
    for (size_t g = 0; g < rows.size(); ++g) {
        std::cout << rows[g].id << '\t' << rows[g].name;
        for (size_t i = 0; i < repGens.size(); ++i) {
            const auto &r = res[i][g];
            if (r.isRepresented()) { std::cout << '\t' << r.getGenomeId(); ++foundCount; } else { std::cout << '\t'; ++errorCount; }
        }
        std::cout << '\n';
    }
    if (a.flag.count("v")) std::cerr << gCount << " genomes processed. " << foundCount << " representatives found; " << errorCount << " failures.\n";
    return 0;
}

// ---- gtoreps (GtoRepGenomeProcessor.java:136-314) --------------------------------------------------------------
// Report the closest representative of every genome of a GTO directory.  With --create / --update an outlier (nothing
// close, or below the threshold) is ADDED to the database and the genomes after it see it (:296-303), which makes the scan
// sequential; it is batched here as: findClosest of all remaining genomes, rows up to and including the first outlier are
// final, add that outlier, repeat from the next genome.  Without update it is one batched findClosest.
// Not reproduced: --save (writing outlier GTOs).
static int cmd_gtoreps(int argc, char **argv) {
    Args a = parse(argc, argv, 2, {{"--create", "create"}, {"--update", "update"}, {"--filter", "filter"}, {"--blackList", "black"}, {"--rna", "rna"},
                                   {"-o", "o"}, {"--output", "o"}, {"-t", "t"}, {"--type", "t"}, {"-v", "v"}},
                   {"create", "update", "filter", "black", "o", "t"});
    if (a.pos.size() != 2) throw std::runtime_error("ParseFailureException: Arguments \"repDb.ser\" and \"gtoDir\" are required");
    const std::string repDbFile = a.pos[0], gtoDir = a.pos[1];
    std::unique_ptr<RepGenomeDb> repDb;
    std::string updateFile = a.opt.count("update") ? a.opt["update"] : std::string();
    if (!a.opt.count("create")) {   // :146-153
        std::ifstream probe(repDbFile);
        if (!probe) throw std::runtime_error("java.io.FileNotFoundException: Representative-genome database " + repDbFile + " not found or unreadable.");
        repDb = RepGenomeDb::load(repDbFile);
    } else if (!updateFile.empty()) {
        throw std::runtime_error("ParseFailureException: Update file is mutually exclusive with create mode.");   // :155
    } else {
        repDb = std::make_unique<RepGenomeDb>(std::stoi(a.opt["create"]));   // :159-160
        updateFile = repDbFile;
    }
    std::vector<std::string> files = genomeDirectory(gtoDir);
    std::map<std::string, std::string> byId;
    for (auto &f : files) { std::string id = f.substr(f.find_last_of('/') + 1); byId[id.substr(0, id.size() - 4)] = f; }
    std::vector<std::string> idList;
    if (a.opt.count("filter")) idList = readFirstColumn(a.opt["filter"]);   // :166
    else for (auto &kv : byId) idList.push_back(kv.first);
    std::set<std::string> blacklist;
    if (a.opt.count("black")) { auto v = readFirstColumn(a.opt["black"]); blacklist.insert(v.begin(), v.end()); }
    const bool rnaFlag = a.flag.count("rna") != 0;
    std::ofstream fout;
    if (a.opt.count("o")) { fout.open(a.opt["o"]); if (!fout) throw std::runtime_error("java.io.FileNotFoundException: " + a.opt["o"]); }
    std::ostream &writer = a.opt.count("o") ? static_cast<std::ostream &>(fout) : std::cout;
    writer << "genome_id\tgenome_name\trep_id\trep_name\tsimilarity\tdistance\toutlier\n";   // :198
    std::vector<RepGenome> todo, deferrals;
    int skipCount = 0;
    for (const std::string &genomeId : idList) {   // :206-241
        auto it = byId.find(genomeId);
        if (it == byId.end()) continue;            // "completely acceptable for the genome to not be in the source"
        Genome genome(it->second);
        std::unique_ptr<RepGenome> rep = getSeedProtein(*repDb, genome);
        if (!rep) { ++skipCount; continue; }
        if (blacklist.count(genome.getId())) { deferrals.push_back(*rep); continue; }
        if (rnaFlag) {
            if (genome.getSsuRRna().empty()) { ++skipCount; continue; }
            if (genome.getSsuRRna().size() < 1400) { deferrals.push_back(*rep); continue; }   // MIN_RNA_LEN :78
        }
        todo.push_back(*rep);
    }
    for (auto &d : deferrals) todo.push_back(d);   // :243-247
    int outCount = 0;
    auto row = [&](const RepGenome &rep, const std::string &repId, const std::string &repName, int sim, double dist, const char *flag) {
        writer << rep.getGenomeId() << '\t' << rep.getName() << '\t' << repId << '\t' << repName << '\t' << sim << '\t'
               << java::format_fixed(dist, 8, 4) << '\t' << flag << '\n';
    };
    for (size_t pos = 0; pos < todo.size();) {   // processGenome :270-314
        std::vector<const ProteinKmers *> qp;
        for (size_t i = pos; i < todo.size(); ++i) qp.push_back(&todo[i]);
        auto res = repDb->findClosest(qp.begin(), qp.end());
        size_t i = pos;
        for (; i < todo.size(); ++i) {
            const auto &r = res[i - pos];
            std::string repId, repName; int sim = 0; double dist = 1.0; bool outlier = false;
            if (r.getSimilarity() == 0) outlier = true;
            else {
                repId = r.getGenomeId(); repName = r.getRepresentative()->getName(); sim = r.getSimilarity(); dist = r.getDistance();
                if (sim < repDb->getThreshold()) outlier = true;
            }
            if (outlier) {
                ++outCount;
                if (!updateFile.empty()) { repDb->addRep(todo[i]); repId = todo[i].getGenomeId(); sim = 9999; dist = 0.0; }
            }
            row(todo[i], repId, repName, sim, dist, outlier ? "*" : "");
            if (outlier && !updateFile.empty()) { ++i; break; }   // the database changed: the rest must be scored again
        }
        pos = i;
    }
    if (a.flag.count("v")) std::cerr << outCount << " outliers found, " << skipCount << " skipped due to missing seed protein.\n";
    if (!updateFile.empty()) repDb->save(updateFile);   // :250-253
    return 0;
}

// ---- buildfromgtos (GtoRepBuildProcessor.java:96-163) ------------------------------------------------------------
// The role definition file (one role; RoleMap.load) is read as: last tab-separated field of its first non-empty line.
static int cmd_buildfromgtos(int argc, char **argv) {
    Args a = parse(argc, argv, 2, {{"-K", "K"}, {"--kmer", "K"}, {"-m", "m"}, {"--sim", "m"}, {"--verify", "verify"}, {"--batch", "batch"}, {"-v", "v"}},
                   {"K", "m", "verify", "batch"});
    if (a.pos.size() != 3) throw std::runtime_error("ParseFailureException: Arguments \"gtoDir\", \"role.definitions\" and \"repDb\" are required");
    const int kmerSize = a.opt.count("K") ? std::stoi(a.opt["K"]) : 8, minSim = a.opt.count("m") ? std::stoi(a.opt["m"]) : 100;   // :78-79
    if (kmerSize < 2) throw std::runtime_error("ParseFailureException: Kmer size must be at least 2.");
    if (minSim < 1) throw std::runtime_error("ParseFailureException: Minimum similarity must be at least 1.");
    uint32_t batch = a.opt.count("batch") ? (uint32_t)std::stoul(a.opt["batch"]) : 0;
    std::ifstream rf(a.pos[1]);
    if (!rf) throw std::runtime_error("java.io.IOException: Role definition file is not found or unreadable.");
    std::vector<std::string> roleLines; std::string line;
    while (std::getline(rf, line)) { if (!line.empty() && line.back() == '\r') line.pop_back(); if (!line.empty()) roleLines.push_back(line); }
    if (roleLines.size() != 1) throw std::runtime_error("ParseFailureException: Role definition file must contain exactly 1 role.");   // :101
    const std::string role = roleLines[0].substr(roleLines[0].find_last_of('\t') == std::string::npos ? 0 : roleLines[0].find_last_of('\t') + 1);
    if (a.opt.count("verify")) { std::ifstream vf(a.opt["verify"]); if (!vf) throw std::runtime_error("java.io.IOException: Verification file is not found or unreadable."); }
    ProteinKmers::setKmerSize(kmerSize);
    RepGenomeDb repDb(minSim, role);
    std::vector<RepGenome> repGenomes;
    for (const std::string &file : genomeDirectory(a.pos[0])) {   // computeRepGenome :170-203
        Genome genome(file);
        const std::string fid = findSeed(genome, role);
        if (fid.empty()) continue;
        repGenomes.emplace_back(fid, genome.getName(), genome.getFeature(fid)->getProteinTranslation());
    }
    repDb.checkGenomes(repGenomes, batch);   // :129-137
    repDb.save(a.pos[2]);
    if (!a.opt.count("verify")) { std::cout << "No verification requested.\n"; return 0; }   // :144
    std::cout << "seq_id\tname\tsimilarity\n";
    std::vector<Sequence> seqs = FastaInputStream(a.opt["verify"]).readAll();
    std::vector<ProteinKmers> qs;
    for (auto &q : seqs) qs.emplace_back(q.sequence);
    std::vector<const ProteinKmers *> qp;
    for (auto &q : qs) qp.push_back(&q);
    auto res = repDb.findClosest(qp.begin(), qp.end());
    for (size_t i = 0; i < seqs.size(); ++i)
        if (!res[i].isRepresented()) std::cout << seqs[i].label << '\t' << seqs[i].comment << '\t' << res[i].getSimilarity() << '\n';
    return 0;
}

// ---- neighbors (NeighborProcessor.java:130-230) ---------------------------------------------------------
// The reference reads genomes from a GenomeSource and takes each one's seed protein and DNA length; GTO ingest is
// outside the path, so the input here is a tab-delimited file with a header line and the columns
// seed_fid, genome_name, dna_size, seed_protein.  Output text is the reference's, row for row.
static int cmd_neighbors(int argc, char **argv) {
    Args a = parse(argc, argv, 2, {{"-n", "n"}, {"--num", "n"}, {"--min", "min"}, {"-v", "v"}}, {"n", "min"});
    if (a.pos.size() != 2) throw std::runtime_error("ParseFailureException: Arguments \"repDbFile\" and \"inFile\" are required");
    int num = a.opt.count("n") ? std::stoi(a.opt["n"]) : 5, minSim = a.opt.count("min") ? std::stoi(a.opt["min"]) : 1;   // :106-109
    if (num <= 0) throw std::runtime_error("ParseFailureException: Close-genome count (num) must be at least 1.");   // :115
    if (minSim < 0) throw std::runtime_error("ParseFailureException: Minimum similarity cannot be zero.");            // :118
    { std::ifstream probe(a.pos[0]); if (!probe) throw std::runtime_error("java.io.FileNotFoundException: Repgen database file " + a.pos[0] + " is not found or unreadable."); }
    auto repDb = RepGenomeDb::load(a.pos[0]);
    std::ifstream in(a.pos[1]);
    if (!in) throw std::runtime_error("java.io.FileNotFoundException: Input genome source " + a.pos[1] + " does not exist.");
    std::vector<RepGenome> seeds; std::vector<long> dnaSize;
    std::string line;
    std::getline(in, line);   // header
    while (std::getline(in, line)) {
        if (!line.empty() && line.back() == '\r') line.pop_back();
        if (line.empty()) continue;
        std::vector<std::string> f; size_t b = 0;
        for (size_t t; (t = line.find('\t', b)) != std::string::npos; b = t + 1) f.push_back(line.substr(b, t - b));
        f.push_back(line.substr(b));
        if (f.size() < 4) throw std::runtime_error("java.io.IOException: input line needs 4 fields: " + line);
        seeds.emplace_back(f[0], f[1], f[3]);
        dnaSize.push_back(std::stol(f[2]));
    }
    std::cout << "genome_id\tgenome_name\tsize\trep_id\trep_name\tsim\tdistance\tclosest\tsignal\n";   // :135
    std::vector<const RepGenome *> reps;
    std::vector<int32_t> sims = repDb->similarities(seeds.begin(), seeds.end(), &reps);
    const size_t R = reps.size();
    for (size_t g = 0; g < seeds.size(); ++g) {   // processGenome :155-203, writeResults :212-230
        const int32_t *row = sims.data() + g * R;
        std::string prefix = seeds[g].getGenomeId() + "\t" + seeds[g].getName() + "\t" + std::to_string(dnaSize[g]);
        auto found = RepGenomeDb::closestN(row, R, num, minSim);
        if (found.empty()) { std::cout << prefix << "\t\t\t\t\t\t\n"; continue; }
        const char *closeFlag = "Y";
        for (size_t r : found) {
            double dist = ProteinKmers::distance(row[r], seeds[g], *reps[r]);
            double signal = (double)dnaSize[g] * (1.0 - dist);
            std::cout << prefix << '\t' << reps[r]->getGenomeId() << '\t' << reps[r]->getName() << '\t' << row[r] << '\t'
                      << java::double_to_string(dist) << '\t' << closeFlag << '\t' << java::double_to_string(signal) << '\n';
            closeFlag = " ";
        }
    }
    return 0;
}

// ---- buildsets: the similarity work of the `genomes` pipeline ------------------------------------------
// Input: one FASTA record per genome (label = seed-protein feature id, comment = genome name, sequence = seed
// protein) in master-list order (ProteinDataFactory orders by rating, score, id; that ordering needs PATRIC
// metadata and is done upstream of this command).  For every threshold, smallest first: the reps of the previous
// set are copied in (copyReps, BaseGenomeProcessor.java:172-176), then every genome is offered in master order
// (buildRepGenSet :184-214).  Outputs per threshold: rep<thr>.ser (saveRepGenSets :433-442) and rep<thr>.list.tbl
// (writeListFiles :348-407) without the three taxonomy columns, which are PATRIC metadata.
static int cmd_buildsets(int argc, char **argv) {
    Args a = parse(argc, argv, 2, {{"-K", "K"}, {"--kmer", "K"}, {"-t", "t"}, {"--repSizes", "t"}, {"-p", "p"}, {"--prot", "p"}, {"--batch", "batch"}, {"-v", "v"}},
                   {"K", "t", "p", "batch"});
    if (a.pos.size() != 2) throw std::runtime_error("ParseFailureException: Arguments \"outDir\" and \"inFile\" are required");
    ProteinKmers::setKmerSize(a.opt.count("K") ? std::stoi(a.opt["K"]) : 8);
    std::string protName = a.opt.count("p") ? a.opt["p"] : RepGenomeDb::DEFAULT_PROTEIN;
    uint32_t batch = a.opt.count("batch") ? (uint32_t)std::stoul(a.opt["batch"]) : 0;
    std::vector<int> sizes;
    {
        std::stringstream ss(a.opt.count("t") ? a.opt["t"] : "10,50,100,200");   // GenomeProcessor.java:62
        std::string tok;
        while (std::getline(ss, tok, ',')) sizes.push_back(std::stoi(tok));
        for (size_t i = 1; i < sizes.size(); ++i)
            if (sizes[i] <= sizes[i - 1]) throw std::runtime_error("ParseFailureException: thresholds must increase");
    }
    const std::string outDir = a.pos[0];
    std::vector<Sequence> seqs = FastaInputStream(a.pos[1]).readAll();
    std::vector<RepGenome> genomes;
    for (auto &s : seqs) genomes.emplace_back(s);
    std::vector<std::unique_ptr<RepGenomeDb>> sets;
    for (int sz : sizes) sets.push_back(std::make_unique<RepGenomeDb>(sz, protName));
    for (size_t k = 0; k < sets.size(); ++k) {   // collateGenomes :151-164
        if (k > 0) for (const RepGenome *r : sets[k - 1]->all()) sets[k]->addRep(*r);
        sets[k]->checkGenomes(genomes, batch);
        if (a.flag.count("v")) std::cerr << sets[k]->toString() << ": " << sets[k]->size() << " representatives.\n";
    }
    std::vector<const ProteinKmers *> qp;
    for (auto &g : genomes) qp.push_back(&g);
    for (auto &db : sets) {
        db->save(outDir + "/rep" + std::to_string(db->getThreshold()) + ".ser");
        std::ofstream lf(outDir + "/" + db->getListFileName());
        if (!lf) throw std::runtime_error("java.io.FileNotFoundException: " + outDir + "/" + db->getListFileName());
        lf << "genome_id\tgenome_name\trep_id\tscore\tdistance\n";
        auto res = db->findClosest(qp.begin(), qp.end());
        for (size_t i = 0; i < genomes.size(); ++i) {
            const auto &r = res[i];
            lf << genomes[i].getGenomeId() << '\t' << genomes[i].getName() << '\t' << (r.getRepresentative() ? r.getGenomeId() : std::string("null")) << '\t'
               << r.getSimilarity() << '\t' << java::format_fixed(r.getDistance(), 4, 2) << '\n';
        }
    }
    return 0;
}

// kmerops: the k-mer set members of SequenceKmers that TargetProcessor.java:160-195 uses (retainAll, removeAll, size, the
// iteration over the k-mers) on the sequences of a FASTA.  Record 0 is the "other" set; for every further record prints
//   label <TAB> size <TAB> |retainAll(other)| <TAB> |removeAll(other)| <TAB> similarity(edited-retain, other) <TAB> the retained k-mers, comma-joined, sorted
static int cmd_kmerops(int argc, char **argv) {
    Args a = parse(argc, argv, 2, {{"-K", "K"}, {"--kmer", "K"}}, {"K"});
    if (a.opt.count("K")) ProteinKmers::setKmerSize(std::stoi(a.opt["K"]));
    FastaInputStream in(std::cin);
    Sequence s;
    if (!in.next(s)) return 0;
    ProteinKmers other(s.sequence);
    while (in.next(s)) {
        ProteinKmers keep(s.sequence), drop(s.sequence), plain(s.sequence);
        keep.retainAll(other);
        drop.removeAll(other);
        std::vector<std::string> k = keep.kmers();
        std::sort(k.begin(), k.end());
        std::cout << s.label << '\t' << plain.size() << '\t' << keep.size() << '\t' << drop.size() << '\t' << keep.similarity(other) << '\t';
        for (size_t i = 0; i < k.size(); ++i) std::cout << (i ? "," : "") << k[i];
        std::cout << '\t' << (keep.equals(plain) ? 1 : 0) << (drop.equals(plain) ? 1 : 0) << '\n';
    }
    return 0;
}

int main(int argc, char **argv) {
    if (argc < 2) { usage(); return 2; }
    std::ios::sync_with_stdio(false);
    try {
        std::string cmd = argv[1];
        if (cmd == "repdb") return cmd_repdb(argc, argv);
        if (cmd == "distances") return cmd_distances(argc, argv);
        if (cmd == "group") return cmd_group(argc, argv);
        if (cmd == "classify") return cmd_classify(argc, argv);
        if (cmd == "gtoclass") return cmd_gtoclass(argc, argv);
        if (cmd == "gtoreps") return cmd_gtoreps(argc, argv);
        if (cmd == "buildfromgtos") return cmd_buildfromgtos(argc, argv);
        if (cmd == "neighbors") return cmd_neighbors(argc, argv);
        if (cmd == "buildsets") return cmd_buildsets(argc, argv);
        if (cmd == "kmerops") return cmd_kmerops(argc, argv);
        usage();
        return 2;
    } catch (const std::exception &e) {   // App.java:129: print the error, exit non-zero
        std::cerr << "Error: " << e.what() << '\n';
        return 1;
    }
}
