// yaml_lite.hpp — the subset of YAML the racer-tracer scene and config files use, parsed into an
// ordered tree.  Block mappings and sequences by indentation, flow mappings `{a: 1}` and flow
// sequences `[1, 2]` (nested, possibly spanning lines), plain / single- / double-quoted scalars,
// comments, a leading `---`.  A key whose value sits alone on the following, deeper-indented line
// (`loader:\n  Sandbox`, racer-tracer/config.yml:28-29) is a scalar.  No anchors, tags or multi-
// document streams: the reference's files (resources/scenes/*.yml, config.yml) do not use them.
//
// Stands where the `config` 0.13 crate stands in the reference (src/scene/yml.rs:154-169,
// src/config.rs:217-224); like that crate, lower_keys() lower-cases every map key.
#pragma once
#include <algorithm>
#include <cctype>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

namespace yaml_lite {

struct ParseError : std::runtime_error {
    using std::runtime_error::runtime_error;
};

struct Node {
    enum Kind { Null, Scalar, Map, Seq } kind = Null;
    std::string scalar;
    bool quoted = false;
    std::vector<std::pair<std::string, Node>> map;   // insertion order
    std::vector<Node> seq;

    bool is_null() const { return kind == Null; }
    bool is_scalar() const { return kind == Scalar; }
    bool is_map() const { return kind == Map; }
    bool is_seq() const { return kind == Seq; }
    const Node* find(const std::string& key) const {
        if (kind != Map) return nullptr;
        for (auto& kv : map)
            if (kv.first == key) return &kv.second;
        return nullptr;
    }
    bool has(const std::string& key) const { const Node* n = find(key); return n && !n->is_null(); }
    const Node& at(const std::string& key) const {
        const Node* n = find(key);
        if (!n) throw ParseError("missing field `" + key + "`");
        return *n;
    }
    double as_double() const {
        if (kind != Scalar) throw ParseError("expected a number");
        char* end = nullptr;
        const double v = std::strtod(scalar.c_str(), &end);
        if (end == scalar.c_str() || *end != '\0') throw ParseError("invalid number `" + scalar + "`");
        return v;
    }
    long long as_int() const { return (long long)as_double(); }
    const std::string& as_string() const {
        if (kind != Scalar) throw ParseError("expected a string");
        return scalar;
    }
    std::vector<std::string> sorted_keys() const {
        std::vector<std::string> k;
        for (auto& kv : map) k.push_back(kv.first);
        std::sort(k.begin(), k.end());
        return k;
    }
};

inline std::string to_lower(std::string s) {
    for (auto& c : s) c = (char)std::tolower((unsigned char)c);
    return s;
}

inline void lower_keys(Node& n) {
    for (auto& kv : n.map) { kv.first = to_lower(kv.first); lower_keys(kv.second); }
    for (auto& v : n.seq) lower_keys(v);
}

namespace detail {

struct Line {
    int indent;
    std::string text;   // without indentation, comment and trailing blanks
    int number;
};

inline std::string strip_comment(const std::string& s) {
    char q = 0;
    for (size_t i = 0; i < s.size(); ++i) {
        const char c = s[i];
        if (q) { if (c == q) q = 0; continue; }
        if (c == '"' || c == '\'') { q = c; continue; }
        if (c == '#' && (i == 0 || s[i - 1] == ' ' || s[i - 1] == '\t')) return s.substr(0, i);
    }
    return s;
}

inline std::string trim(const std::string& s) {
    size_t a = 0, b = s.size();
    while (a < b && (s[a] == ' ' || s[a] == '\t' || s[a] == '\r')) ++a;
    while (b > a && (s[b - 1] == ' ' || s[b - 1] == '\t' || s[b - 1] == '\r')) --b;
    return s.substr(a, b - a);
}

inline Node scalar_node(std::string s) {
    Node n;
    s = trim(s);
    if (s.size() >= 2 && ((s.front() == '"' && s.back() == '"') || (s.front() == '\'' && s.back() == '\''))) {
        n.kind = Node::Scalar; n.quoted = true; n.scalar = s.substr(1, s.size() - 2);
        return n;
    }
    if (s.empty() || s == "~" || s == "null" || s == "Null" || s == "NULL") return n;   // Null
    n.kind = Node::Scalar; n.scalar = s;
    return n;
}

// position of the ':' that ends a plain or quoted key ("key: value" / "key:"), or npos
inline size_t key_colon(const std::string& s) {
    char q = 0;
    int depth = 0;
    for (size_t i = 0; i < s.size(); ++i) {
        const char c = s[i];
        if (q) { if (c == q) q = 0; continue; }
        if (c == '"' || c == '\'') { q = c; continue; }
        if (c == '[' || c == '{') ++depth;
        if (c == ']' || c == '}') --depth;
        if (c == ':' && depth == 0 && (i + 1 == s.size() || s[i + 1] == ' ' || s[i + 1] == '\t')) return i;
    }
    return std::string::npos;
}

struct Flow {   // flow-style value parser over one string
    const std::string& s;
    size_t i = 0;
    explicit Flow(const std::string& str) : s(str) {}
    void ws() { while (i < s.size() && (s[i] == ' ' || s[i] == '\t' || s[i] == '\n')) ++i; }
    Node value() {
        ws();
        if (i >= s.size()) return Node();
        if (s[i] == '[') return sequence();
        if (s[i] == '{') return mapping();
        return scalar(",]}");
    }
    Node scalar(const char* stops) {
        ws();
        size_t a = i;
        if (i < s.size() && (s[i] == '"' || s[i] == '\'')) {
            const char q = s[i++];
            while (i < s.size() && s[i] != q) ++i;
            if (i >= s.size()) throw ParseError("unterminated quoted string");
            ++i;
            return scalar_node(s.substr(a, i - a));
        }
        while (i < s.size() && !std::strchr(stops, s[i])) {
            if (s[i] == ':' && (i + 1 >= s.size() || s[i + 1] == ' ') && std::strchr(stops, ':')) break;
            ++i;
        }
        return scalar_node(s.substr(a, i - a));
    }
    Node sequence() {
        Node n; n.kind = Node::Seq;
        ++i;
        for (;;) {
            ws();
            if (i >= s.size()) throw ParseError("unterminated flow sequence");
            if (s[i] == ']') { ++i; return n; }
            n.seq.push_back(value());
            ws();
            if (i < s.size() && s[i] == ',') ++i;
        }
    }
    Node mapping() {
        Node n; n.kind = Node::Map;
        ++i;
        for (;;) {
            ws();
            if (i >= s.size()) throw ParseError("unterminated flow mapping");
            if (s[i] == '}') { ++i; return n; }
            Node k = scalar(",}:");
            ws();
            Node v;
            if (i < s.size() && s[i] == ':') { ++i; v = value(); }
            n.map.emplace_back(k.kind == Node::Scalar ? k.scalar : std::string(), v);
            ws();
            if (i < s.size() && s[i] == ',') ++i;
        }
    }
};

struct Parser {
    std::vector<Line> lines;
    size_t pos = 0;

    [[noreturn]] void fail(const std::string& what) const {
        const int ln = pos < lines.size() ? lines[pos].number : (lines.empty() ? 0 : lines.back().number);
        throw ParseError(what + " at line " + std::to_string(ln));
    }

    // an inline value that starts on this line; flow collections may continue on following lines
    Node inline_value(std::string text) {
        text = trim(text);
        if (!text.empty() && (text[0] == '[' || text[0] == '{')) {
            auto balance = [](const std::string& t) {
                int d = 0; char q = 0;
                for (char c : t) {
                    if (q) { if (c == q) q = 0; continue; }
                    if (c == '"' || c == '\'') q = c;
                    else if (c == '[' || c == '{') ++d;
                    else if (c == ']' || c == '}') --d;
                }
                return d;
            };
            while (balance(text) > 0 && pos < lines.size()) text += "\n" + lines[pos++].text;
            try {
                Flow f(text);
                return f.value();
            } catch (const ParseError& e) { fail(e.what()); }
        }
        return scalar_node(text);
    }

    Node block(int indent) {
        if (pos >= lines.size() || lines[pos].indent < indent) return Node();
        const int ind = lines[pos].indent;
        const std::string& first = lines[pos].text;
        if (first == "-" || first.rfind("- ", 0) == 0) return sequence(ind);
        if (key_colon(first) == std::string::npos) {   // a lone scalar (value of the key above)
            Node n = inline_value(lines[pos++].text);
            return n;
        }
        return mapping(ind);
    }

    Node mapping(int ind) {
        Node n; n.kind = Node::Map;
        while (pos < lines.size() && lines[pos].indent == ind) {
            const std::string text = lines[pos].text;
            if (text == "-" || text.rfind("- ", 0) == 0) break;
            const size_t c = key_colon(text);
            if (c == std::string::npos) fail("expected `key: value`");
            Node k = scalar_node(text.substr(0, c));
            const std::string rest = trim(text.substr(c + 1));
            ++pos;
            Node v;
            if (!rest.empty()) v = inline_value(rest);
            else if (pos < lines.size() && lines[pos].indent > ind) v = block(lines[pos].indent);
            else if (pos < lines.size() && lines[pos].indent == ind &&
                     (lines[pos].text == "-" || lines[pos].text.rfind("- ", 0) == 0)) v = sequence(ind);   // "key:\n- a"
            n.map.emplace_back(k.kind == Node::Scalar ? k.scalar : std::string(), v);
        }
        if (pos < lines.size() && lines[pos].indent > ind) fail("unexpected indentation");
        return n;
    }

    Node sequence(int ind) {
        Node n; n.kind = Node::Seq;
        while (pos < lines.size() && lines[pos].indent == ind &&
               (lines[pos].text == "-" || lines[pos].text.rfind("- ", 0) == 0)) {
            std::string rest = lines[pos].text.size() > 1 ? lines[pos].text.substr(2) : std::string();
            const int inner = ind + 2 + (int)(rest.size() - trim(rest).size() > 0 ? rest.find_first_not_of(' ') : 0);
            rest = trim(rest);
            if (rest.empty()) {
                ++pos;
                n.seq.push_back(pos < lines.size() && lines[pos].indent > ind ? block(lines[pos].indent) : Node());
            } else if (key_colon(rest) != std::string::npos && rest[0] != '[' && rest[0] != '{') {
                // "- key: value" opens a mapping whose further keys are indented under the key
                lines[pos].indent = inner;
                lines[pos].text = rest;
                n.seq.push_back(mapping(inner));
            } else {
                ++pos;
                n.seq.push_back(inline_value(rest));
            }
        }
        return n;
    }
};

}  // namespace detail

inline Node parse(const std::string& text) {
    detail::Parser p;
    size_t a = 0;
    int number = 0;
    while (a <= text.size()) {
        size_t b = text.find('\n', a);
        if (b == std::string::npos) b = text.size();
        std::string raw = text.substr(a, b - a);
        a = b + 1;
        ++number;
        raw = detail::strip_comment(raw);
        int indent = 0;
        while ((size_t)indent < raw.size() && raw[indent] == ' ') ++indent;
        if ((size_t)indent < raw.size() && raw[indent] == '\t') throw ParseError("tab indentation at line " + std::to_string(number));
        const std::string t = detail::trim(raw);
        if (t.empty() || t == "---" || t == "...") continue;
        p.lines.push_back({indent, t, number});
    }
    if (p.lines.empty()) return Node();
    Node root = p.block(p.lines[0].indent);
    if (p.pos < p.lines.size()) p.fail("unexpected content");
    return root;
}

}  // namespace yaml_lite
