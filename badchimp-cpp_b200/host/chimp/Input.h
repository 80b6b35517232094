// Minimal reader for the reference's input deck syntax (src/io/Input.h:24-71): blocks
//   <name>  key v1 v2 ...  <end>
// plus top-level "key value" lines and '#' comments; input["fluid"]["tau"] converts to int, double,
// std::string or std::vector<double>.  Only what the three target mains read is covered (no inline
// math, no $variables) -- the config layer is host-only and not part of the hot path.
#ifndef CHIMP_INPUT_H
#define CHIMP_INPUT_H

#include <fstream>
#include <map>
#include <sstream>

#include "LBglobal.h"

class Block
{
public:
    Block &operator[](const std::string &key)
    {
        auto it = child_.find(key);
        if (it == child_.end()) chimp_host::die("Input: keyword '" + key + "' not found");
        return it->second;
    }
    Block &operator[](const char *key) { return (*this)[std::string(key)]; }
    Block &operator[](int i) { idx_ = i; return *this; }
    operator double() { const double v = std::stod(values_.at(idx_)); idx_ = 0; return v; }
    operator int() { const int v = int(std::stod(values_.at(idx_))); idx_ = 0; return v; }
    operator std::string() const { return values_.at(0); }
    operator std::vector<double>() const
    {
        std::vector<double> v;
        for (const auto &s : values_) v.push_back(std::stod(s));
        return v;
    }
    std::map<std::string, Block> child_;
    std::vector<std::string> values_;

private:
    int idx_ = 0;
};

class Input : public Block
{
public:
    explicit Input(const std::string &filename)
    {
        std::ifstream in(filename);
        if (!in) chimp_host::die("Input: could not open " + filename);
        std::string line;
        Block *cur = this;
        while (std::getline(in, line)) {
            const auto hash = line.find('#');
            if (hash != std::string::npos) line.erase(hash);
            std::istringstream ss(line);
            std::string key, tok;
            if (!(ss >> key)) continue;
            if (key == "<end>") { cur = this; continue; }
            if (key.front() == '<' && key.back() == '>') { cur = &child_[key.substr(1, key.size() - 2)]; continue; }
            Block &b = cur->child_[key];
            while (ss >> tok) b.values_.push_back(tok);
        }
    }
};

#endif
