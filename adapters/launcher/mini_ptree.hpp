/* mini_ptree.hpp -- the subset of boost::property_tree::ptree + read_json that slam_launcher.cpp uses
 * (slam_launcher.cpp:905-925 and every Create* factory above it), for building the launcher where Boost
 * is absent.  Same access semantics:
 *   get<T>(path)            value at a dotted path, throws std::runtime_error if the key is missing
 *   get(path, default)      value or default; "true" / "false" strings convert to bool like Boost does
 *   get_child(path)         subtree, throws if missing;  get_child_optional(path): testable + dereferenceable
 *   begin() / end()         children as (key, subtree) pairs; JSON array elements have empty keys
 *   get_value<T>()          the node's own value
 *   put(path, value)        create / overwrite (used for --set overrides on the command line)
 * A tree with Boost uses boost::property_tree::ptree instead; the factory templates in
 * lgs_adapters/create_cuda_backends.hpp accept either. */
#ifndef LGS_LAUNCHER_MINI_PTREE_HPP
#define LGS_LAUNCHER_MINI_PTREE_HPP

#include <cctype>
#include <cstdlib>
#include <fstream>
#include <sstream>
#include <stdexcept>
#include <string>
#include <type_traits>
#include <utility>
#include <vector>

namespace LgsLauncher {

class Ptree
{
public:
    using Child = std::pair<std::string, Ptree>;
    using const_iterator = std::vector<Child>::const_iterator;

    /* Optional subtree: what get_child_optional returns */
    class Optional
    {
    public:
        explicit Optional(const Ptree* p = nullptr) : mPtr(p) { }
        explicit operator bool() const { return this->mPtr != nullptr; }
        const Ptree& operator*() const { return *this->mPtr; }
        const Ptree* operator->() const { return this->mPtr; }
    private:
        const Ptree* mPtr;
    };

    const_iterator begin() const { return this->mChildren.begin(); }
    const_iterator end() const { return this->mChildren.end(); }
    bool empty() const { return this->mChildren.empty(); }
    const std::string& data() const { return this->mValue; }

    template <typename T> T get_value() const { return Convert<T>(this->mValue, "<value>"); }

    Optional get_child_optional(const std::string& path) const { return Optional(this->Find(path)); }

    const Ptree& get_child(const std::string& path) const
    {
        const Ptree* p = this->Find(path);
        if (p == nullptr)
            throw std::runtime_error("settings: no such group: " + path);
        return *p;
    }

    template <typename T> T get(const std::string& path) const
    {
        const Ptree* p = this->Find(path);
        if (p == nullptr)
            throw std::runtime_error("settings: no such key: " + path);
        return Convert<T>(p->mValue, path);
    }

    template <typename T> T get(const std::string& path, const T& defaultValue) const
    {
        const Ptree* p = this->Find(path);
        return p == nullptr ? defaultValue : Convert<T>(p->mValue, path);
    }

    std::string get(const std::string& path, const char* defaultValue) const
    { return this->get<std::string>(path, std::string(defaultValue)); }

    void put(const std::string& path, const std::string& value)
    {
        Ptree* node = this;
        std::size_t pos = 0;
        while (pos <= path.size()) {
            const std::size_t dot = path.find('.', pos);
            const std::string key = path.substr(pos, dot == std::string::npos ? std::string::npos : dot - pos);
            Ptree* next = nullptr;
            for (auto& c : node->mChildren)
                if (c.first == key) { next = &c.second; break; }
            if (next == nullptr) {
                node->mChildren.emplace_back(key, Ptree());
                next = &node->mChildren.back().second;
            }
            node = next;
            if (dot == std::string::npos)
                break;
            pos = dot + 1;
        }
        node->mValue = value;
        node->mChildren.clear();
    }

    /* Parse a JSON document (objects, arrays, strings, numbers, true / false / null) */
    static Ptree ParseJson(const std::string& text)
    {
        std::size_t pos = 0;
        Ptree root = ParseValue(text, pos);
        SkipSpace(text, pos);
        if (pos != text.size())
            throw std::runtime_error("settings: trailing characters after the JSON document");
        return root;
    }

private:
    std::string        mValue;
    std::vector<Child> mChildren;

    const Ptree* Find(const std::string& path) const
    {
        const Ptree* node = this;
        std::size_t pos = 0;
        while (true) {
            const std::size_t dot = path.find('.', pos);
            const std::string key = path.substr(pos, dot == std::string::npos ? std::string::npos : dot - pos);
            const Ptree* next = nullptr;
            for (const auto& c : node->mChildren)
                if (c.first == key) { next = &c.second; break; }
            if (next == nullptr)
                return nullptr;
            node = next;
            if (dot == std::string::npos)
                return node;
            pos = dot + 1;
        }
    }

    template <typename T> static T Convert(const std::string& v, const std::string& what)
    {
        if constexpr (std::is_same<T, std::string>::value) {
            return v;
        } else if constexpr (std::is_same<T, bool>::value) {
            if (v == "true" || v == "1") return true;
            if (v == "false" || v == "0") return false;
            throw std::runtime_error("settings: " + what + " is not a boolean: " + v);
        } else {
            char* end = nullptr;
            const double d = std::strtod(v.c_str(), &end);
            if (end == v.c_str() || *end != '\0')
                throw std::runtime_error("settings: " + what + " is not a number: " + v);
            return static_cast<T>(d);
        }
    }

    static void SkipSpace(const std::string& s, std::size_t& pos)
    { while (pos < s.size() && std::isspace(static_cast<unsigned char>(s[pos]))) ++pos; }

    static std::string ParseString(const std::string& s, std::size_t& pos)
    {
        std::string out;
        ++pos;                                               /* opening quote */
        while (pos < s.size() && s[pos] != '"') {
            if (s[pos] == '\\' && pos + 1 < s.size()) {
                const char e = s[++pos];
                out += e == 'n' ? '\n' : e == 't' ? '\t' : e;
            } else {
                out += s[pos];
            }
            ++pos;
        }
        if (pos >= s.size())
            throw std::runtime_error("settings: unterminated string");
        ++pos;                                               /* closing quote */
        return out;
    }

    static Ptree ParseValue(const std::string& s, std::size_t& pos)
    {
        SkipSpace(s, pos);
        if (pos >= s.size())
            throw std::runtime_error("settings: unexpected end of the JSON document");
        Ptree node;
        if (s[pos] == '{') {
            ++pos;
            SkipSpace(s, pos);
            if (pos < s.size() && s[pos] == '}') { ++pos; return node; }
            while (true) {
                SkipSpace(s, pos);
                if (pos >= s.size() || s[pos] != '"')
                    throw std::runtime_error("settings: expected a key at offset " + std::to_string(pos));
                std::string key = ParseString(s, pos);
                SkipSpace(s, pos);
                if (pos >= s.size() || s[pos] != ':')
                    throw std::runtime_error("settings: expected ':' after key " + key);
                ++pos;
                node.mChildren.emplace_back(std::move(key), ParseValue(s, pos));
                SkipSpace(s, pos);
                if (pos < s.size() && s[pos] == ',') { ++pos; continue; }
                if (pos < s.size() && s[pos] == '}') { ++pos; return node; }
                throw std::runtime_error("settings: expected ',' or '}' at offset " + std::to_string(pos));
            }
        }
        if (s[pos] == '[') {
            ++pos;
            SkipSpace(s, pos);
            if (pos < s.size() && s[pos] == ']') { ++pos; return node; }
            while (true) {
                node.mChildren.emplace_back(std::string(), ParseValue(s, pos));
                SkipSpace(s, pos);
                if (pos < s.size() && s[pos] == ',') { ++pos; continue; }
                if (pos < s.size() && s[pos] == ']') { ++pos; return node; }
                throw std::runtime_error("settings: expected ',' or ']' at offset " + std::to_string(pos));
            }
        }
        if (s[pos] == '"') {
            node.mValue = ParseString(s, pos);
            return node;
        }
        const std::size_t start = pos;                       /* number, true, false, null */
        while (pos < s.size() && s[pos] != ',' && s[pos] != '}' && s[pos] != ']' &&
               !std::isspace(static_cast<unsigned char>(s[pos])))
            ++pos;
        node.mValue = s.substr(start, pos - start);
        if (node.mValue == "null")
            node.mValue.clear();
        return node;
    }
};

inline void ReadJson(const std::string& fileName, Ptree& tree)
{
    std::ifstream file(fileName);
    if (!file)
        throw std::runtime_error("settings: cannot open " + fileName);
    std::stringstream buffer;
    buffer << file.rdbuf();
    tree = Ptree::ParseJson(buffer.str());
}

} /* namespace LgsLauncher */

#endif /* LGS_LAUNCHER_MINI_PTREE_HPP */
