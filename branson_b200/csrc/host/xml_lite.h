// xml_lite.h -- minimal XML element tree for Branson decks.
//
// The reference parses its decks with the vendored pugixml DOM (src/input.h:72-102) and only ever reads element
// text (child("name").text().as_double() / child_value("name")).  The schema has no attributes, entities or CDATA,
// so a ~150-line recursive-descent reader is enough; numbers go through strtod / strtoll exactly as pugixml's
// as_double / as_int do (leading blanks and forms like ".65" occur in the reference decks).
#pragma once
#include <cctype>
#include <cstdlib>
#include <fstream>
#include <memory>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

namespace xml_lite {

struct Node {
  std::string name;
  std::string text;  // concatenated character data directly inside this element
  std::vector<std::unique_ptr<Node>> children;

  const Node *child(const std::string &n) const {
    for (const auto &c : children)
      if (c->name == n) return c.get();
    return nullptr;
  }
  // pugixml child_value(): raw text of the named child (not trimmed, like pugixml's default parse flags), "" if absent
  std::string child_value(const std::string &n) const {
    const Node *c = child(n);
    return c ? c->text : std::string();
  }
  std::string trimmed() const {
    size_t b = 0, e = text.size();
    while (b < e && std::isspace((unsigned char)text[b])) ++b;
    while (e > b && std::isspace((unsigned char)text[e - 1])) --e;
    return text.substr(b, e - b);
  }
  // pugixml text().as_double(): strtod of the text, 0 if the node is missing
  double as_double(const std::string &n, double def = 0.0) const {
    const Node *c = child(n);
    return c ? std::strtod(c->text.c_str(), nullptr) : def;
  }
  long long as_llong(const std::string &n, long long def = 0) const {
    const Node *c = child(n);
    if (!c) return def;
    const char *s = c->text.c_str();
    while (*s && std::isspace((unsigned char)*s)) ++s;
    const char *p = (*s == '-' || *s == '+') ? s + 1 : s;
    const int base = (p[0] == '0' && (p[1] == 'x' || p[1] == 'X')) ? 16 : 10;
    return std::strtoll(s, nullptr, base);
  }
  int as_int(const std::string &n, int def = 0) const { return (int)as_llong(n, def); }
};

class Parser {
public:
  explicit Parser(const std::string &s) : src(s), pos(0) {}

  std::unique_ptr<Node> parse_document() {
    auto root = std::make_unique<Node>();
    root->name = "#document";
    for (;;) {
      skip_misc();
      if (pos >= src.size()) break;
      if (src[pos] != '<') throw std::runtime_error("xml: text outside of the root element");
      root->children.push_back(parse_element());
    }
    return root;
  }

private:
  const std::string &src;
  size_t pos;

  bool starts(const char *lit) const { return src.compare(pos, std::char_traits<char>::length(lit), lit) == 0; }
  void skip_until(const char *lit) {
    const size_t p = src.find(lit, pos);
    if (p == std::string::npos) throw std::runtime_error(std::string("xml: missing ") + lit);
    pos = p + std::char_traits<char>::length(lit);
  }
  void skip_misc() {
    for (;;) {
      while (pos < src.size() && std::isspace((unsigned char)src[pos])) ++pos;
      if (starts("<!--")) skip_until("-->");
      else if (starts("<?")) skip_until("?>");
      else if (starts("<!")) skip_until(">");
      else return;
    }
  }
  std::unique_ptr<Node> parse_element() {
    ++pos;  // '<'
    auto node = std::make_unique<Node>();
    while (pos < src.size() && !std::isspace((unsigned char)src[pos]) && src[pos] != '>' && src[pos] != '/')
      node->name += src[pos++];
    if (node->name.empty()) throw std::runtime_error("xml: empty tag name");
    // attributes are not part of the deck schema: skip to the end of the tag
    bool self_closing = false;
    while (pos < src.size() && src[pos] != '>') {
      if (src[pos] == '"' || src[pos] == '\'') {
        const char q = src[pos++];
        while (pos < src.size() && src[pos] != q) ++pos;
      }
      self_closing = (src[pos] == '/');
      ++pos;
    }
    if (pos >= src.size()) throw std::runtime_error("xml: unterminated tag <" + node->name);
    ++pos;  // '>'
    if (self_closing) return node;
    for (;;) {
      if (pos >= src.size()) throw std::runtime_error("xml: missing </" + node->name + ">");
      if (starts("<!--")) { skip_until("-->"); continue; }
      if (starts("<?")) { skip_until("?>"); continue; }
      if (starts("</")) {
        pos += 2;
        std::string close;
        while (pos < src.size() && src[pos] != '>') close += src[pos++];
        ++pos;
        while (!close.empty() && std::isspace((unsigned char)close.back())) close.pop_back();
        if (close != node->name) throw std::runtime_error("xml: </" + close + "> closes <" + node->name + ">");
        return node;
      }
      if (src[pos] == '<') { node->children.push_back(parse_element()); continue; }
      node->text += src[pos++];
    }
  }
};

inline std::unique_ptr<Node> parse_file(const std::string &path) {
  std::ifstream in(path);
  if (!in) throw std::runtime_error("cannot open " + path);
  std::stringstream ss;
  ss << in.rdbuf();
  const std::string s = ss.str();
  Parser p(s);
  return p.parse_document();
}

}  // namespace xml_lite
