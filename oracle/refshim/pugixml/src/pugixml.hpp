// Minimal pugixml-API-compatible DOM used ONLY to build the reference oracle
// (oracle/_ref) from the unmodified sources under /root/reference.
//
// The reference's lib/pugixml submodule directory is empty in this checkout, so
// the reference cannot be compiled as shipped.  This header is an independent,
// from-scratch implementation of the small slice of the pugixml interface the
// reference touches (see SURVEY.md section 7 step 0 for the list).  It is test
// infrastructure: nothing in the product (quickrank_b200/, host/) includes it.
//
// None of the hot-path arithmetic goes through this file: it only formats and
// parses model files.
#ifndef QRB200_ORACLE_PUGIXML_STANDIN_HPP
#define QRB200_ORACLE_PUGIXML_STANDIN_HPP

#include <cstddef>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <sstream>
#include <string>
#include <vector>

namespace pugi {

typedef char char_t;

const unsigned int format_indent = 0x01;
const unsigned int format_no_declaration = 0x08;
const unsigned int format_default = format_indent;

enum xml_node_type { node_null, node_document, node_element, node_pcdata };

namespace impl {
struct attr_rec {
  std::string name, value;
};
struct node_rec {
  xml_node_type type = node_element;
  std::string name;    // element name
  std::string value;   // pcdata payload (type == node_pcdata)
  node_rec *parent = nullptr;
  std::vector<node_rec *> kids;
  std::vector<attr_rec *> attrs;
  ~node_rec() {
    for (auto *k : kids) delete k;
    for (auto *a : attrs) delete a;
  }
};
inline std::string fmt_double(double v) {
  char buf[64];
  snprintf(buf, sizeof(buf), "%.17g", v);
  return buf;
}
inline std::string fmt_float(float v) {
  char buf[64];
  snprintf(buf, sizeof(buf), "%.9g", (double) v);
  return buf;
}
inline void escape(std::ostream &os, const std::string &s, bool attr) {
  for (char c : s) {
    switch (c) {
      case '&': os << "&amp;"; break;
      case '<': os << "&lt;"; break;
      case '>': os << "&gt;"; break;
      case '"': if (attr) os << "&quot;"; else os << c; break;
      default: os << c;
    }
  }
}
inline std::string unescape(const std::string &s) {
  std::string o;
  o.reserve(s.size());
  for (size_t i = 0; i < s.size(); ++i) {
    if (s[i] == '&') {
      if (!s.compare(i, 5, "&amp;")) { o += '&'; i += 4; continue; }
      if (!s.compare(i, 4, "&lt;")) { o += '<'; i += 3; continue; }
      if (!s.compare(i, 4, "&gt;")) { o += '>'; i += 3; continue; }
      if (!s.compare(i, 6, "&quot;")) { o += '"'; i += 5; continue; }
      if (!s.compare(i, 6, "&apos;")) { o += '\''; i += 5; continue; }
    }
    o += s[i];
  }
  return o;
}
}  // namespace impl

class xml_attribute {
 public:
  xml_attribute() : a_(nullptr) {}
  explicit xml_attribute(impl::attr_rec *a) : a_(a) {}
  bool empty() const { return !a_; }
  explicit operator bool() const { return a_ != nullptr; }
  bool operator!() const { return !a_; }
  const char_t *name() const { return a_ ? a_->name.c_str() : ""; }
  const char_t *value() const { return a_ ? a_->value.c_str() : ""; }
  const char_t *as_string(const char_t *def = "") const {
    return a_ ? a_->value.c_str() : def;
  }
  int as_int(int def = 0) const { return a_ ? (int) strtol(value(), 0, 10) : def; }
  unsigned int as_uint(unsigned int def = 0) const {
    return a_ ? (unsigned int) strtoul(value(), 0, 10) : def;
  }
  double as_double(double def = 0) const { return a_ ? strtod(value(), 0) : def; }
  float as_float(float def = 0) const { return a_ ? (float) strtod(value(), 0) : def; }
  bool as_bool(bool def = false) const {
    if (!a_ || a_->value.empty()) return def;
    char c = a_->value[0];
    return c == '1' || c == 't' || c == 'T' || c == 'y' || c == 'Y';
  }
  bool set_value(const char_t *v) { if (!a_) return false; a_->value = v; return true; }
  xml_attribute &operator=(const char_t *v) { set_value(v); return *this; }
  xml_attribute &operator=(const std::string &v) { set_value(v.c_str()); return *this; }
  xml_attribute &operator=(int v) { set_value(std::to_string(v).c_str()); return *this; }
  xml_attribute &operator=(unsigned int v) { set_value(std::to_string(v).c_str()); return *this; }
  xml_attribute &operator=(long v) { set_value(std::to_string(v).c_str()); return *this; }
  xml_attribute &operator=(unsigned long v) { set_value(std::to_string(v).c_str()); return *this; }
  xml_attribute &operator=(long long v) { set_value(std::to_string(v).c_str()); return *this; }
  xml_attribute &operator=(unsigned long long v) { set_value(std::to_string(v).c_str()); return *this; }
  xml_attribute &operator=(double v) { set_value(impl::fmt_double(v).c_str()); return *this; }
  xml_attribute &operator=(float v) { set_value(impl::fmt_float(v).c_str()); return *this; }
  xml_attribute &operator=(bool v) { set_value(v ? "true" : "false"); return *this; }

 private:
  impl::attr_rec *a_;
};

class xml_node;
class xml_node_range;

class xml_text {
 public:
  xml_text() : n_(nullptr) {}
  explicit xml_text(impl::node_rec *n) : n_(n) {}
  bool empty() const { return !find(); }
  explicit operator bool() const { return find() != nullptr; }
  const char_t *get() const { auto *d = find(); return d ? d->value.c_str() : ""; }
  const char_t *as_string(const char_t *def = "") const {
    auto *d = find(); return d ? d->value.c_str() : def;
  }
  int as_int(int def = 0) const { auto *d = find(); return d ? (int) strtol(d->value.c_str(), 0, 10) : def; }
  unsigned int as_uint(unsigned int def = 0) const {
    auto *d = find(); return d ? (unsigned int) strtoul(d->value.c_str(), 0, 10) : def;
  }
  double as_double(double def = 0) const { auto *d = find(); return d ? strtod(d->value.c_str(), 0) : def; }
  float as_float(float def = 0) const { auto *d = find(); return d ? (float) strtod(d->value.c_str(), 0) : def; }
  bool as_bool(bool def = false) const {
    auto *d = find();
    if (!d || d->value.empty()) return def;
    char c = d->value[0];
    return c == '1' || c == 't' || c == 'T' || c == 'y' || c == 'Y';
  }
  bool set(const char_t *v) {
    if (!n_) return false;
    auto *d = find();
    if (!d) {
      d = new impl::node_rec();
      d->type = node_pcdata;
      d->parent = n_;
      n_->kids.insert(n_->kids.begin(), d);
    }
    d->value = v;
    return true;
  }
  xml_text &operator=(const char_t *v) { set(v); return *this; }
  xml_text &operator=(const std::string &v) { set(v.c_str()); return *this; }
  xml_text &operator=(int v) { set(std::to_string(v).c_str()); return *this; }
  xml_text &operator=(unsigned int v) { set(std::to_string(v).c_str()); return *this; }
  xml_text &operator=(long v) { set(std::to_string(v).c_str()); return *this; }
  xml_text &operator=(unsigned long v) { set(std::to_string(v).c_str()); return *this; }
  xml_text &operator=(long long v) { set(std::to_string(v).c_str()); return *this; }
  xml_text &operator=(unsigned long long v) { set(std::to_string(v).c_str()); return *this; }
  xml_text &operator=(double v) { set(impl::fmt_double(v).c_str()); return *this; }
  xml_text &operator=(float v) { set(impl::fmt_float(v).c_str()); return *this; }
  xml_text &operator=(bool v) { set(v ? "true" : "false"); return *this; }

 private:
  impl::node_rec *find() const {
    if (!n_) return nullptr;
    if (n_->type == node_pcdata) return n_;
    for (auto *k : n_->kids)
      if (k->type == node_pcdata) return k;
    return nullptr;
  }
  impl::node_rec *n_;
};

class xml_node {
 public:
  xml_node() : n_(nullptr) {}
  explicit xml_node(impl::node_rec *n) : n_(n) {}
  bool empty() const { return !n_; }
  explicit operator bool() const { return n_ != nullptr; }
  bool operator!() const { return !n_; }
  bool operator==(const xml_node &o) const { return n_ == o.n_; }
  bool operator!=(const xml_node &o) const { return n_ != o.n_; }
  xml_node_type type() const { return n_ ? n_->type : node_null; }
  const char_t *name() const { return n_ ? n_->name.c_str() : ""; }
  const char_t *value() const { return n_ ? n_->value.c_str() : ""; }
  bool set_name(const char_t *nm) { if (!n_) return false; n_->name = nm; return true; }
  xml_node parent() const { return xml_node(n_ ? n_->parent : nullptr); }

  xml_node child(const char_t *nm) const {
    if (n_)
      for (auto *k : n_->kids)
        if (k->type == node_element && k->name == nm) return xml_node(k);
    return xml_node();
  }
  xml_node first_child() const {
    return xml_node(n_ && !n_->kids.empty() ? n_->kids.front() : nullptr);
  }
  xml_node next_sibling() const {
    if (!n_ || !n_->parent) return xml_node();
    auto &v = n_->parent->kids;
    for (size_t i = 0; i + 1 < v.size(); ++i)
      if (v[i] == n_) return xml_node(v[i + 1]);
    return xml_node();
  }
  xml_attribute attribute(const char_t *nm) const {
    if (n_)
      for (auto *a : n_->attrs)
        if (a->name == nm) return xml_attribute(a);
    return xml_attribute();
  }
  xml_text text() const { return xml_text(n_); }
  const char_t *child_value() const {
    if (n_)
      for (auto *k : n_->kids)
        if (k->type == node_pcdata) return k->value.c_str();
    return "";
  }
  const char_t *child_value(const char_t *nm) const { return child(nm).child_value(); }

  xml_node append_child(const char_t *nm) {
    if (!n_) return xml_node();
    auto *k = new impl::node_rec();
    k->type = node_element;
    k->name = nm;
    k->parent = n_;
    n_->kids.push_back(k);
    return xml_node(k);
  }
  xml_attribute append_attribute(const char_t *nm) {
    if (!n_) return xml_attribute();
    auto *a = new impl::attr_rec();
    a->name = nm;
    n_->attrs.push_back(a);
    return xml_attribute(a);
  }
  xml_node append_copy(const xml_node &proto) {
    if (!n_ || !proto.n_) return xml_node();
    auto *k = clone(proto.n_);
    k->parent = n_;
    n_->kids.push_back(k);
    return xml_node(k);
  }
  bool remove_child(const xml_node &c) {
    if (!n_ || !c.n_) return false;
    for (size_t i = 0; i < n_->kids.size(); ++i)
      if (n_->kids[i] == c.n_) {
        delete n_->kids[i];
        n_->kids.erase(n_->kids.begin() + i);
        return true;
      }
    return false;
  }
  bool remove_child(const char_t *nm) { return remove_child(child(nm)); }

  // range support: children() and children(name) (defined after the class)
  inline xml_node_range children() const;
  inline xml_node_range children(const char_t *nm) const;

  // Absolute, attribute-free location paths only ("/a/b/c"): all the reference uses.
  class xpath_node_stub {
   public:
    xpath_node_stub() {}
    explicit xpath_node_stub(impl::node_rec *n) : n_(n) {}
    xml_node node() const { return xml_node(n_); }
   private:
    impl::node_rec *n_ = nullptr;
  };
  typedef std::vector<xpath_node_stub> xpath_node_set_stub;
  xpath_node_set_stub select_nodes(const char_t *path) const {
    std::vector<impl::node_rec *> cur;
    impl::node_rec *r = n_;
    while (r && r->parent) r = r->parent;
    if (r) cur.push_back(r);
    std::stringstream ss(path);
    std::string seg;
    while (std::getline(ss, seg, '/')) {
      if (seg.empty()) continue;
      std::vector<impl::node_rec *> nxt;
      for (auto *c : cur)
        for (auto *k : c->kids)
          if (k->type == node_element && k->name == seg) nxt.push_back(k);
      cur.swap(nxt);
    }
    xpath_node_set_stub out;
    for (auto *c : cur) out.push_back(xpath_node_stub(c));
    return out;
  }
  xpath_node_stub select_node(const char_t *path) const {
    auto s = select_nodes(path);
    return s.empty() ? xpath_node_stub() : s[0];
  }

  void print(std::ostream &os, const char_t *indent = "\t", unsigned int flags = format_default,
             int depth = 0) const {
    if (n_) write(os, n_, indent, (flags & format_indent) != 0, depth);
  }

 protected:
  static impl::node_rec *clone(const impl::node_rec *s) {
    auto *d = new impl::node_rec();
    d->type = s->type;
    d->name = s->name;
    d->value = s->value;
    for (auto *a : s->attrs) d->attrs.push_back(new impl::attr_rec(*a));
    for (auto *k : s->kids) {
      auto *c = clone(k);
      c->parent = d;
      d->kids.push_back(c);
    }
    return d;
  }
  static void write(std::ostream &os, const impl::node_rec *n, const char *indent, bool pretty,
                    int depth) {
    if (n->type == node_document) {
      for (auto *k : n->kids) write(os, k, indent, pretty, depth);
      return;
    }
    if (n->type == node_pcdata) {
      impl::escape(os, n->value, false);
      return;
    }
    if (pretty) for (int i = 0; i < depth; ++i) os << indent;
    os << '<' << n->name;
    for (auto *a : n->attrs) {
      os << ' ' << a->name << "=\"";
      impl::escape(os, a->value, true);
      os << '"';
    }
    if (n->kids.empty()) {
      os << " />";
      if (pretty) os << '\n';
      return;
    }
    os << '>';
    bool only_text = true;
    for (auto *k : n->kids) if (k->type != node_pcdata) only_text = false;
    if (only_text) {
      for (auto *k : n->kids) impl::escape(os, k->value, false);
    } else {
      if (pretty) os << '\n';
      for (auto *k : n->kids) {
        if (k->type == node_pcdata) {
          if (pretty) for (int i = 0; i <= depth; ++i) os << indent;
          impl::escape(os, k->value, false);
          if (pretty) os << '\n';
        } else {
          write(os, k, indent, pretty, depth + 1);
        }
      }
      if (pretty) for (int i = 0; i < depth; ++i) os << indent;
    }
    os << "</" << n->name << '>';
    if (pretty) os << '\n';
  }
  impl::node_rec *n_;
};


class xml_node_range {
 public:
  class iterator {
   public:
    iterator(const std::vector<impl::node_rec *> *v, size_t i, const char *filt)
        : v_(v), i_(i), filt_(filt) { skip(); cur_ = at(); }
    // iterator traits and a non-const dereference, as pugixml's xml_node_iterator has them
    typedef std::ptrdiff_t difference_type;
    typedef xml_node value_type;
    typedef xml_node *pointer;
    typedef xml_node &reference;
    typedef std::forward_iterator_tag iterator_category;
    xml_node &operator*() const { return cur_; }
    xml_node *operator->() const { return &cur_; }
    iterator operator++(int) { iterator t = *this; ++*this; return t; }
    iterator &operator++() { ++i_; skip(); cur_ = at(); return *this; }
    bool operator!=(const iterator &o) const { return i_ != o.i_; }
    bool operator==(const iterator &o) const { return i_ == o.i_; }
   private:
    void skip() {
      while (v_ && i_ < v_->size() &&
             ((*v_)[i_]->type != node_element ||
              (filt_ && (*v_)[i_]->name != filt_)))
        ++i_;
    }
    xml_node at() const {
      return (v_ && i_ < v_->size()) ? xml_node((*v_)[i_]) : xml_node();
    }
    const std::vector<impl::node_rec *> *v_;
    size_t i_;
    const char *filt_;
    mutable xml_node cur_;
  };
  xml_node_range(const std::vector<impl::node_rec *> *v, const char *filt)
      : v_(v), filt_(filt ? filt : ""), has_filt_(filt != nullptr) {}
  iterator begin() const { return iterator(v_, 0, has_filt_ ? filt_.c_str() : nullptr); }
  iterator end() const { return iterator(v_, v_ ? v_->size() : 0, nullptr); }
 private:
  const std::vector<impl::node_rec *> *v_;
  std::string filt_;
  bool has_filt_;
};
inline xml_node_range xml_node::children() const {
  return xml_node_range(n_ ? &n_->kids : nullptr, nullptr);
}
inline xml_node_range xml_node::children(const char_t *nm) const {
  return xml_node_range(n_ ? &n_->kids : nullptr, nm);
}

typedef xml_node::xpath_node_stub xpath_node;
typedef xml_node::xpath_node_set_stub xpath_node_set;

struct xml_parse_result {
  bool ok = false;
  std::string msg;
  operator bool() const { return ok; }
  const char *description() const { return msg.c_str(); }
};

class xml_document : public xml_node {
 public:
  xml_document() { reset(); }
  ~xml_document() { delete n_; }
  xml_document(const xml_document &) = delete;
  xml_document &operator=(const xml_document &) = delete;
  void reset() {
    delete n_;
    n_ = new impl::node_rec();
    n_->type = node_document;
  }
  xml_node document_element() const {
    for (auto *k : n_->kids) if (k->type == node_element) return xml_node(k);
    return xml_node();
  }
  xml_parse_result load_string(const char_t *s) {
    reset();
    return parse(std::string(s));
  }
  xml_parse_result load_file(const char *path) {
    reset();
    std::ifstream f(path, std::ios::binary);
    xml_parse_result r;
    if (!f) { r.msg = "File was not found"; return r; }
    std::stringstream ss;
    ss << f.rdbuf();
    return parse(ss.str());
  }
  void save(std::ostream &os, const char_t *indent = "\t", unsigned int flags = format_default) const {
    if (!(flags & format_no_declaration)) {
      os << "<?xml version=\"1.0\"?>";
      if (flags & format_indent) os << '\n';
    }
    print(os, indent, flags, 0);
  }
  bool save_file(const char *path, const char_t *indent = "\t",
                 unsigned int flags = format_default) const {
    std::ofstream f(path, std::ios::binary);
    if (!f) return false;
    save(f, indent, flags);
    return (bool) f;
  }

 private:
  static bool is_ws(char c) { return c == ' ' || c == '\t' || c == '\n' || c == '\r'; }
  xml_parse_result parse(const std::string &s) {
    xml_parse_result r;
    impl::node_rec *cur = n_;
    size_t i = 0, n = s.size();
    while (i < n) {
      if (s[i] == '<') {
        if (!s.compare(i, 4, "<!--")) {
          size_t e = s.find("-->", i + 4);
          if (e == std::string::npos) { r.msg = "unterminated comment"; return r; }
          i = e + 3;
        } else if (!s.compare(i, 2, "<?")) {
          size_t e = s.find("?>", i + 2);
          if (e == std::string::npos) { r.msg = "unterminated declaration"; return r; }
          i = e + 2;
        } else if (!s.compare(i, 9, "<![CDATA[")) {
          size_t e = s.find("]]>", i + 9);
          if (e == std::string::npos) { r.msg = "unterminated cdata"; return r; }
          auto *d = new impl::node_rec();
          d->type = node_pcdata;
          d->value = s.substr(i + 9, e - i - 9);
          d->parent = cur;
          cur->kids.push_back(d);
          i = e + 3;
        } else if (!s.compare(i, 2, "<!")) {
          size_t e = s.find('>', i);
          if (e == std::string::npos) { r.msg = "unterminated doctype"; return r; }
          i = e + 1;
        } else if (i + 1 < n && s[i + 1] == '/') {
          size_t e = s.find('>', i);
          if (e == std::string::npos) { r.msg = "unterminated end tag"; return r; }
          std::string nm = s.substr(i + 2, e - i - 2);
          while (!nm.empty() && is_ws(nm.back())) nm.pop_back();
          if (cur == n_ || cur->name != nm) { r.msg = "end tag mismatch"; return r; }
          cur = cur->parent;
          i = e + 1;
        } else {
          size_t j = i + 1;
          while (j < n && !is_ws(s[j]) && s[j] != '>' && s[j] != '/') ++j;
          auto *el = new impl::node_rec();
          el->type = node_element;
          el->name = s.substr(i + 1, j - i - 1);
          el->parent = cur;
          cur->kids.push_back(el);
          bool selfclose = false;
          for (;;) {
            while (j < n && is_ws(s[j])) ++j;
            if (j >= n) { r.msg = "unterminated start tag"; return r; }
            if (s[j] == '>') { ++j; break; }
            if (s[j] == '/') { selfclose = true; ++j; continue; }
            size_t k = j;
            while (k < n && s[k] != '=' && !is_ws(s[k]) && s[k] != '>') ++k;
            auto *a = new impl::attr_rec();
            a->name = s.substr(j, k - j);
            el->attrs.push_back(a);
            while (k < n && is_ws(s[k])) ++k;
            if (k < n && s[k] == '=') {
              ++k;
              while (k < n && is_ws(s[k])) ++k;
              if (k >= n || (s[k] != '"' && s[k] != '\'')) { r.msg = "bad attribute"; return r; }
              char q = s[k];
              size_t e = s.find(q, k + 1);
              if (e == std::string::npos) { r.msg = "bad attribute"; return r; }
              a->value = impl::unescape(s.substr(k + 1, e - k - 1));
              k = e + 1;
            }
            j = k;
          }
          if (!selfclose) cur = el;
          i = j;
        }
      } else {
        size_t e = s.find('<', i);
        if (e == std::string::npos) e = n;
        std::string t = s.substr(i, e - i);
        bool allws = true;
        for (char c : t) if (!is_ws(c)) { allws = false; break; }
        if (!allws && cur != n_) {
          auto *d = new impl::node_rec();
          d->type = node_pcdata;
          d->value = impl::unescape(t);
          d->parent = cur;
          cur->kids.push_back(d);
        }
        i = e;
      }
    }
    if (cur != n_) { r.msg = "unexpected end of data"; return r; }
    r.ok = true;
    r.msg = "No error";
    return r;
  }
};

}  // namespace pugi

#endif
