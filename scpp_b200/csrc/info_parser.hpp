// scpp_b200/csrc/info_parser.hpp — minimal reader for the Boost.PropertyTree INFO files the reference uses
// (scpp_core/utils/include/parameterServer.hpp:34-127; Boost is absent from this image).  Supports what the shipped
// configs contain: `key value`, `key { (i) value ... }` blocks with optional `scaling`, `;` comments.
// Semantics mirrored: loadScalar throws on a missing key (:66-77); loadMatrix throws on missing / redundant entries
// (:95-103) and applies `scaling` (:84,126); the constructor only REPORTS an unreadable file (:39-48) — here it throws,
// since silently continuing with an empty tree would only fail later at the first loadScalar.
#pragma once
#include <cmath>
#include <fstream>
#include <map>
#include <memory>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

namespace scpp {

class ParameterServer {
public:
    explicit ParameterServer(const std::string &filename)
    {
        std::ifstream f(filename);
        if (!f) throw std::runtime_error("Could not open file for reading: " + filename);
        std::vector<std::string> tok;
        std::string line;
        while (std::getline(f, line)) {
            const size_t c = line.find(';');
            if (c != std::string::npos) line.erase(c);
            std::string cur;
            for (char ch : line) {
                if (ch == '{' || ch == '}') { if (!cur.empty()) { tok.push_back(cur); cur.clear(); } tok.push_back(std::string(1, ch)); }
                else if (ch == ' ' || ch == '\t' || ch == '\r') { if (!cur.empty()) { tok.push_back(cur); cur.clear(); } }
                else cur += ch;
            }
            if (!cur.empty()) tok.push_back(cur);
            tok.push_back("\n");
        }
        size_t i = 0;
        parse(tok, i, root_, 0);
    }

    template <typename T>
    void loadScalar(const std::string &name, T &out) const
    {
        auto it = root_.find(name);
        if (it == root_.end() || it->second.block) throw std::runtime_error("WARNING: Failed to load scalar type: " + name + "!\n");
        out = convert<T>(it->second.value, name);
    }
    // vector of fixed length n: entries (0)..(n-1)
    void loadMatrix(const std::string &name, double *v, int n) const
    {
        auto it = root_.find(name);
        if (it == root_.end() || !it->second.block) throw std::runtime_error("Failed to load matrix type: " + name + "!\n");
        const auto &blk = *it->second.block;
        double scaling = 1.;
        auto sc = blk.find("scaling");
        if (sc != blk.end()) scaling = convert<double>(sc->second.value, name);
        const int entries = int(blk.size()) - (sc != blk.end() ? 1 : 0);
        if (entries < n) throw std::runtime_error("Missing entries in matrix type: " + name + "!\n");
        if (entries > n) throw std::runtime_error("Redundant entries in matrix type: " + name + "!\n");
        for (int i = 0; i < n; i++) {
            auto e = blk.find("(" + std::to_string(i) + ")");
            if (e == blk.end()) throw std::runtime_error("Failed to load matrix type: " + name + "!\n");
            v[i] = convert<double>(e->second.value, name) * scaling;
        }
    }

private:
    struct Node { std::string value; std::shared_ptr<std::map<std::string, Node>> block; };
    std::map<std::string, Node> root_;

    static void parse(const std::vector<std::string> &t, size_t &i, std::map<std::string, Node> &out, int depth)
    {
        while (i < t.size()) {
            if (t[i] == "\n") { i++; continue; }
            if (t[i] == "}") { if (depth == 0) throw std::runtime_error("INFO: unmatched '}'"); i++; return; }
            const std::string key = t[i++];
            Node nd;
            if (i < t.size() && t[i] != "\n" && t[i] != "{" && t[i] != "}") nd.value = t[i++];
            size_t j = i;
            while (j < t.size() && t[j] == "\n") j++;
            if (j < t.size() && t[j] == "{") {
                i = j + 1;
                nd.block = std::make_shared<std::map<std::string, Node>>();
                parse(t, i, *nd.block, depth + 1);
            }
            out[key] = nd;
        }
        if (depth != 0) throw std::runtime_error("INFO: missing '}'");
    }
    template <typename T>
    static T convert(const std::string &s, const std::string &name)
    {
        if constexpr (std::is_same<T, bool>::value) {
            if (s == "true" || s == "1") return true;
            if (s == "false" || s == "0") return false;
            throw std::runtime_error("WARNING: Failed to load scalar type: " + name + "!\n");
        } else {
            std::istringstream is(s);
            T v;
            is >> v;
            if (is.fail() || s.empty()) throw std::runtime_error("WARNING: Failed to load scalar type: " + name + "!\n");
            return v;
        }
    }
};

} // namespace scpp
