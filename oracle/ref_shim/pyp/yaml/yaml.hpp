// ORACLE — TEST INFRASTRUCTURE ONLY. The subset of the `pyp` YAML wrapper the reference's Parse functions use
// (ORBExtractor.cpp:10-18, src/Frame.cpp:343-364, modules/camera/Pinhole.cpp:31-43): a flat `key: value` file, values scalar,
// quoted path, or a whitespace-separated number list; `${CURRENT_FOLDER}` expands to the file's directory (config/frame.yaml:6-8).
#pragma once
#include <fstream>
#include <map>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>
namespace Yaml {
class Node {
   public:
    Node() {}
    explicit Node(const std::string& v, const std::string& dir) : value_(v), dir_(dir) {}
    Node& operator[](const std::string& key) {
        auto it = kids_.find(key);
        if (it == kids_.end()) throw std::runtime_error("yaml: missing key " + key);
        return it->second;
    }
    template <typename T> T As() const { std::istringstream s(value_); T v{}; s >> v; return v; }
    template <typename T> std::vector<T> AsVector() const { std::istringstream s(value_); std::vector<T> v; T x; while (s >> x) v.push_back(x); return v; }
    std::string AsPath() const {
        std::string v = value_;
        const std::string tok = "${CURRENT_FOLDER}";
        size_t p = v.find(tok);
        if (p != std::string::npos) v.replace(p, tok.size(), dir_);
        return v;
    }
    std::map<std::string, Node> kids_;
    std::string value_, dir_;
};
template <> inline std::string Node::As<std::string>() const { return value_; }
static inline void Parse(Node& root, const std::string& path) {
    std::ifstream f(path);
    if (!f) throw std::runtime_error("yaml: cannot open " + path);
    std::string dir = ".";
    size_t sl = path.find_last_of('/');
    if (sl != std::string::npos) dir = path.substr(0, sl);
    std::string line;
    while (std::getline(f, line)) {
        size_t h = line.find('#');
        if (h != std::string::npos) line = line.substr(0, h);
        size_t c = line.find(':');
        if (c == std::string::npos) continue;
        auto trim = [](std::string s) { size_t a = s.find_first_not_of(" \t\r\""), b = s.find_last_not_of(" \t\r\""); return a == std::string::npos ? std::string() : s.substr(a, b - a + 1); };
        std::string k = trim(line.substr(0, c)), v = trim(line.substr(c + 1));
        if (!k.empty()) root.kids_[k] = Node(v, dir);
    }
}
}  // namespace Yaml
