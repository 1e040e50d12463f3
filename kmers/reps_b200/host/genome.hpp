// genome.hpp -- the little of org.theseed.genome.Genome (un-vendored) that the path's callers need: reading a GTO file
// (the JSON "genome typed object" of data/*.gto) and finding the genome's seed protein.  Host-only: file ingest is the
// step before the GPU path, used by the gtoclass command (GtoClassProcessor.java:114-151), which is
// BASELINE configs[0].
#pragma once
#include <dirent.h>

#include "theseed.hpp"

namespace theseed {

// ---- a small JSON reader (objects, arrays, strings with escapes, numbers, literals) ---------------------------------
struct Json {
    enum Type { Null, Bool, Number, String, Array, Object } type = Null;
    bool b = false; double num = 0; std::string str;
    std::vector<Json> arr; std::vector<std::pair<std::string, Json>> obj;
    const Json *get(const std::string &key) const { for (auto &kv : obj) if (kv.first == key) return &kv.second; return nullptr; }
    std::string text(const std::string &key) const { const Json *v = get(key); return v && v->type == String ? v->str : std::string(); }
};

class JsonParser {
  public:
    explicit JsonParser(const std::string &s) : s_(s) {}
    Json parse() { Json v = value(); ws(); if (p_ != s_.size()) fail("trailing characters"); return v; }
  private:
    [[noreturn]] void fail(const char *what) const { throw std::runtime_error(std::string("java.io.IOException: invalid JSON (") + what + ") at offset " + std::to_string(p_)); }
    void ws() { while (p_ < s_.size() && (s_[p_] == ' ' || s_[p_] == '\n' || s_[p_] == '\t' || s_[p_] == '\r')) ++p_; }
    bool eat(char c) { ws(); if (p_ < s_.size() && s_[p_] == c) { ++p_; return true; } return false; }
    static void utf8(std::string &out, unsigned cp) {
        if (cp < 0x80) out += (char)cp;
        else if (cp < 0x800) { out += (char)(0xC0 | (cp >> 6)); out += (char)(0x80 | (cp & 63)); }
        else if (cp < 0x10000) { out += (char)(0xE0 | (cp >> 12)); out += (char)(0x80 | ((cp >> 6) & 63)); out += (char)(0x80 | (cp & 63)); }
        else { out += (char)(0xF0 | (cp >> 18)); out += (char)(0x80 | ((cp >> 12) & 63)); out += (char)(0x80 | ((cp >> 6) & 63)); out += (char)(0x80 | (cp & 63)); }
    }
    unsigned hex4() {
        if (p_ + 4 > s_.size()) fail("short \\u escape");
        unsigned v = 0;
        for (int i = 0; i < 4; ++i) {
            const char c = s_[p_++];
            v = v * 16 + (c >= '0' && c <= '9' ? c - '0' : c >= 'a' && c <= 'f' ? c - 'a' + 10 : c >= 'A' && c <= 'F' ? c - 'A' + 10 : (fail("bad \\u escape"), 0));
        }
        return v;
    }
    std::string string() {
        std::string out;
        for (;;) {
            if (p_ >= s_.size()) fail("unterminated string");
            const char c = s_[p_++];
            if (c == '"') return out;
            if (c != '\\') { out += c; continue; }
            if (p_ >= s_.size()) fail("unterminated escape");
            const char e = s_[p_++];
            switch (e) {
            case '"': out += '"'; break; case '\\': out += '\\'; break; case '/': out += '/'; break;
            case 'b': out += '\b'; break; case 'f': out += '\f'; break; case 'n': out += '\n'; break; case 'r': out += '\r'; break; case 't': out += '\t'; break;
            case 'u': {
                unsigned cp = hex4();
                if (cp >= 0xD800 && cp < 0xDC00 && p_ + 1 < s_.size() && s_[p_] == '\\' && s_[p_ + 1] == 'u') { p_ += 2; const unsigned lo = hex4(); cp = 0x10000 + ((cp - 0xD800) << 10) + (lo - 0xDC00); }
                utf8(out, cp);
                break;
            }
            default: fail("unknown escape");
            }
        }
    }
    Json value() {
        ws();
        if (p_ >= s_.size()) fail("unexpected end");
        Json v;
        const char c = s_[p_];
        if (c == '{') {
            ++p_; v.type = Json::Object;
            if (eat('}')) return v;
            do { ws(); if (p_ >= s_.size() || s_[p_] != '"') fail("object key expected"); ++p_; std::string k = string(); if (!eat(':')) fail("':' expected"); v.obj.emplace_back(std::move(k), value()); } while (eat(','));
            if (!eat('}')) fail("'}' expected");
        } else if (c == '[') {
            ++p_; v.type = Json::Array;
            if (eat(']')) return v;
            do v.arr.push_back(value()); while (eat(','));
            if (!eat(']')) fail("']' expected");
        } else if (c == '"') { ++p_; v.type = Json::String; v.str = string(); }
        else if (!s_.compare(p_, 4, "true")) { p_ += 4; v.type = Json::Bool; v.b = true; }
        else if (!s_.compare(p_, 5, "false")) { p_ += 5; v.type = Json::Bool; }
        else if (!s_.compare(p_, 4, "null")) { p_ += 4; }
        else {
            const size_t b = p_;
            while (p_ < s_.size() && (std::isdigit((unsigned char)s_[p_]) || s_[p_] == '-' || s_[p_] == '+' || s_[p_] == '.' || s_[p_] == 'e' || s_[p_] == 'E')) ++p_;
            if (p_ == b) fail("value expected");
            v.type = Json::Number; v.num = std::strtod(s_.substr(b, p_ - b).c_str(), nullptr);
        }
        return v;
    }
    const std::string &s_; size_t p_ = 0;
};

// ---- Genome / Feature -------------------------------------------------------------------------------------------------
struct Feature {
    std::string id, type, function, proteinTranslation;
    const std::string &getId() const { return id; }
    const std::string &getFunction() const { return function; }
    const std::string &getProteinTranslation() const { return proteinTranslation; }
};

class Genome {
  public:
    explicit Genome(const std::string &file) {   // new Genome(File)
        std::ifstream in(file, std::ios::binary);
        if (!in) throw std::runtime_error("java.io.FileNotFoundException: " + file);
        std::stringstream ss; ss << in.rdbuf();
        const std::string text = ss.str();
        const Json g = JsonParser(text).parse();
        if (g.type != Json::Object) throw std::runtime_error("java.io.IOException: " + file + " is not a genome typed object");
        id_ = g.text("id"); name_ = g.text("scientific_name"); domain_ = g.text("domain"); ssuRRna_ = g.text("ssu_rrna");
        if (const Json *fs = g.get("features"))
            for (const Json &f : fs->arr) features_.push_back(Feature{f.text("id"), f.text("type"), f.text("function"), f.text("protein_translation")});
    }
    const std::string &getId() const { return id_; }
    const std::string &getName() const { return name_; }
    const std::string &getDomain() const { return domain_; }
    const std::string &getSsuRRna() const { return ssuRRna_; }   // GtoRepGenomeProcessor.java:224
    const std::vector<Feature> &getFeatures() const { return features_; }
    const Feature *getFeature(const std::string &fid) const { for (auto &f : features_) if (f.id == fid) return &f; return nullptr; }
    std::string toString() const { return id_ + " (" + name_ + ")"; }
  private:
    std::string id_, name_, domain_, ssuRRna_;
    std::vector<Feature> features_;
};

// RoleUtilities.commentFree: the function without a trailing "# comment" / "! comment"
inline std::string commentFree(const std::string &fn) {
    size_t cut = fn.find_first_of("#!");
    std::string s = cut == std::string::npos ? fn : fn.substr(0, cut);
    while (!s.empty() && std::isspace((unsigned char)s.back())) s.pop_back();
    size_t b = 0; while (b < s.size() && std::isspace((unsigned char)s[b])) ++b;
    return s.substr(b);
}

// GenomeDescriptor.findSeed(genome) (GtoClassProcessor.java:122; un-vendored): the id of the genome's seed protein -- the
// LONGEST feature whose comment-free function is the seed role, with or without its EC number
// (ScoreCouplingReporter.java:135,164-168; TestRolePatterns.java:39-41) -- or "" when there is none
inline std::string findSeed(const Genome &genome, const std::string &role = RepGenomeDb::DEFAULT_PROTEIN) {
    const Feature *best = nullptr;
    for (const Feature &f : genome.getFeatures()) {
        const std::string fn = commentFree(f.function);
        if (fn != role && fn != role + " (EC 6.1.1.20)") continue;
        if (!best || f.proteinTranslation.size() > best->proteinTranslation.size()) best = &f;
    }
    return best ? best->id : std::string();
}

// RepGenomeDb.getSeedProtein(genome) (GtoRepGenomeProcessor.java:213): the genome's seed protein as a RepGenome, found by
// the database's protein name; empty when the genome has none.  (Same rule as GtoRepBuildProcessor.computeRepGenome
// :176-192: the longest protein feature with the role.)
inline std::unique_ptr<RepGenome> getSeedProtein(const RepGenomeDb &db, const Genome &genome) {
    const std::string fid = findSeed(genome, db.getProtName());
    if (fid.empty()) return nullptr;
    return std::make_unique<RepGenome>(fid, genome.getName(), genome.getFeature(fid)->getProteinTranslation());
}

// first column of a tab-delimited file with a header line (TabbedLineReader.readColumn / readSet(file, "1"))
inline std::vector<std::string> readFirstColumn(const std::string &file) {
    std::ifstream in(file);
    if (!in) throw std::runtime_error("java.io.FileNotFoundException: " + file);
    std::vector<std::string> out; std::string line;
    std::getline(in, line);
    while (std::getline(in, line)) {
        if (!line.empty() && line.back() == '\r') line.pop_back();
        if (!line.empty()) out.push_back(line.substr(0, line.find('\t')));
    }
    return out;
}

// GenomeSource of type DIR: the *.gto files of a directory, in genome-id order [recalled: GenomeDirectory keeps a sorted set]
inline std::vector<std::string> genomeDirectory(const std::string &dir) {
    DIR *d = opendir(dir.c_str());
    if (!d) throw std::runtime_error("java.io.FileNotFoundException: Input genome source " + dir + " is not found.");
    std::vector<std::string> ids;
    while (dirent *e = readdir(d)) { std::string n = e->d_name; if (n.size() > 4 && n.compare(n.size() - 4, 4, ".gto") == 0) ids.push_back(n.substr(0, n.size() - 4)); }
    closedir(d);
    std::sort(ids.begin(), ids.end());
    for (auto &i : ids) i = dir + "/" + i + ".gto";
    return ids;
}

}  // namespace theseed
