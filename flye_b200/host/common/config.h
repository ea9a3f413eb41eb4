// flye_b200 host mirror — configuration singletons with the reference's interface
// (reference: src/common/config.h:27-115).  Values are floats; files are "key = value" lines with '#'
// comments and "%include other.cfg" relative to the including file.
#pragma once
#include <cstdlib>
#include <fstream>
#include <sstream>
#include <stdexcept>
#include <string>
#include <unordered_map>

class Config {
public:
    static void load(const std::string& filename) {
        std::ifstream in(filename);
        if (!in) throw std::runtime_error("Can't open config file: " + filename);
        const size_t slash = filename.find_last_of("/\\");
        const std::string dir = slash == std::string::npos ? std::string() : filename.substr(0, slash + 1);
        for (std::string line; std::getline(in, line);) {
            if (line.empty() || line[0] == '#') continue;
            if (line.compare(0, 8, "%include") == 0) {
                std::istringstream ss(line);
                std::string directive, target;
                ss >> directive >> target;
                load(dir + target);
                continue;
            }
            const size_t eq = line.find('=');
            if (eq == std::string::npos || line.find('=', eq + 1) != std::string::npos)
                throw std::runtime_error("Error parsing config file");
            table()[strip(line.substr(0, eq))] = (float)std::atof(strip(line.substr(eq + 1)).c_str());
        }
    }
    static float get(const std::string& key) {
        auto it = table().find(key);
        if (it == table().end()) throw std::runtime_error("No such parameter: " + key);
        return it->second;
    }
    static void addParameters(const std::string& list) {   // "k1=v1,k2=v2"
        std::istringstream ss(list);
        for (std::string item; std::getline(ss, item, ',');) {
            const size_t eq = item.find('=');
            if (eq == std::string::npos) continue;
            table()[strip(item.substr(0, eq))] = (float)std::atof(strip(item.substr(eq + 1)).c_str());
        }
    }

private:
    static std::unordered_map<std::string, float>& table() {
        static std::unordered_map<std::string, float> t;
        return t;
    }
    static std::string strip(const std::string& s) {
        const size_t a = s.find_first_not_of(" \t\r\n"), b = s.find_last_not_of(" \t\r\n");
        return a == std::string::npos ? std::string() : s.substr(a, b - a + 1);
    }
};

struct Parameters {
    static Parameters& get() {
        static Parameters p;
        return p;
    }
    int minimumOverlap = 0;
    size_t kmerSize = 0;
    size_t numThreads = 1;
    bool unevenCoverage = false;
};
