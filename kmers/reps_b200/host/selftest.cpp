// selftest.cpp -- CPU-only checks of the host layer's pure-host parts (no GPU call is made): the
// java.util.HashMap iteration-order emulation, Java number formatting, FASTA I/O and the RepGenomeDb
// load -> save round trip.  Driven by tests/test_host_cpu.py, which compares the output with the oracle's
// independent implementations.
//
//   selftest order              < keys (one per line)        -> iteration order (insertion indices)
//   selftest order-remove [every] < keys, -key, ~key           -> order after puts / Iterator.remove / Map.remove, + invariants
//   selftest fixed W P          < doubles (one per line)     -> String.format("%W.Pf")
//   selftest dtos               < doubles                    -> Double.toString
//   selftest resave IN OUT                                   -> RepGenomeDb::load(IN)->save(OUT)
//   selftest fasta              < FASTA                      -> "label\tcomment\tsequence" lines
//   selftest idcmp              < "a\tb" lines                -> RepGenome.compareTo sign, lexical and natural id order
//   selftest seed FILE.gto ...                               -> GTO ingest + seed-protein lookup, one line per genome
#include "genome.hpp"

using namespace theseed;

int main(int argc, char **argv) {
    if (argc < 2) return 2;
    std::string cmd = argv[1];
    std::ios::sync_with_stdio(false);
    try {
        if (cmd == "order") {
            java::HashMapOrder m;
            std::string line;
            while (std::getline(std::cin, line)) m.put(line);
            for (int i : m.order()) std::cout << i << '\n';
        } else if (cmd == "order-remove") {   // keys; lines starting with '-' remove that key
            java::HashMapOrder m;
            std::string line;
            // '-key': Iterator.remove();  '~key': Map.remove(key);  last line: invariants and number of tree bins
            bool ok = true;
            while (std::getline(std::cin, line)) {
                if (!line.empty() && line[0] == '-') m.remove(line.substr(1), false);
                else if (!line.empty() && line[0] == '~') m.remove(line.substr(1), true);
                else m.put(line);
                if (argc > 2) ok = ok && m.checkInvariants();   // "order-remove every": check after every operation
            }
            for (int i : m.order()) std::cout << i << '\n';
            std::cout << "inv=" << (ok && m.checkInvariants()) << " treebins=" << m.treeBins() << '\n';
        } else if (cmd == "fixed" && argc == 4) {
            int w = std::atoi(argv[2]), p = std::atoi(argv[3]);
            std::string line;
            while (std::getline(std::cin, line)) std::cout << java::format_fixed(std::strtod(line.c_str(), nullptr), w, p) << '\n';
        } else if (cmd == "dtos") {
            std::string line;
            while (std::getline(std::cin, line)) std::cout << java::double_to_string(std::strtod(line.c_str(), nullptr)) << '\n';
        } else if (cmd == "resave" && argc == 4) {
            auto db = RepGenomeDb::load(argv[2]);
            db->save(argv[3]);
            std::cout << db->size() << ' ' << db->getThreshold() << ' ' << db->getKmerSize() << ' ' << db->getProtName() << '\n';
        } else if (cmd == "fasta") {
            FastaInputStream in(std::cin);
            Sequence s;
            while (in.next(s)) std::cout << s.label << '\t' << s.comment << '\t' << s.sequence << '\n';
        } else if (cmd == "seed") {   // GTO files -> "genome id, domain, seed feature id, genome name, seed protein" (no GPU call)
            for (int i = 2; i < argc; ++i) {
                Genome g(argv[i]);
                const std::string fid = findSeed(g);
                std::cout << g.getId() << '\t' << g.getDomain() << '\t' << fid << '\t' << g.getName() << '\t'
                          << (fid.empty() ? std::string() : g.getFeature(fid)->getProteinTranslation()) << '\n';
            }
        } else if (cmd == "idcmp") {     // "a\tb" per line -> RepGenome.compareTo sign under the lexical and the natural id order
            std::string line;
            while (std::getline(std::cin, line)) {
                const size_t t = line.find('\t');
                const std::string a = line.substr(0, t), b = line.substr(t + 1);
                RepGenome::setIdOrder(RepGenome::IdOrder::LEXICAL);
                const int l = RepGenome::compareIds(a, b);
                RepGenome::setIdOrder(RepGenome::IdOrder::NATURAL);
                std::cout << l << '\t' << RepGenome::compareIds(a, b) << '\n';
            }
        } else if (cmd == "genomeof") {
            std::string line;
            while (std::getline(std::cin, line)) std::cout << genomeOf(line) << '\n';
        } else {
            return 2;
        }
    } catch (const std::exception &e) {
        std::cerr << "Error: " << e.what() << '\n';
        return 1;
    }
    return 0;
}
