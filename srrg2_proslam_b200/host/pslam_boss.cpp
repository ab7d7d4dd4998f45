// pslam_boss.cpp -- .conf (BOSS) reader / writer, class registry.  Grammar: SURVEY.md App. C, as used by
// configurations/kitti.conf, icl.conf, euroc.conf of the reference:
//   file   := { '"ClassName"' '{' pair { ',' pair } '}' }
//   pair   := '"key"' ':' value
//   value  := number | '"string"' | '[' [ value { ',' value } ] ']' | '{' '"#pointer"' ':' int '}'
//   `//` starts a comment that runs to the end of the line; objects may be referenced before they are defined.
#include "pslam_boss.hpp"

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <set>

namespace pslam_host {

PropertyBase::PropertyBase(const char* name, const char* description, Configurable* owner, bool* changed_flag)
  : _name(name), _description(description), _changed_flag(changed_flag) {
  owner->_properties.push_back(this);
}

PropertyBase* Configurable::property(const std::string& name) const {
  for (PropertyBase* p : _properties)
    if (p->name() == name) return p;
  return nullptr;
}

ClassRegistry& ClassRegistry::instance() {
  static ClassRegistry r;
  return r;
}

ConfigurablePtr ClassRegistry::create(const std::string& class_name) const {
  auto it = _factories.find(class_name);
  ConfigurablePtr c = it == _factories.end() ? std::make_shared<GenericConfigurable>() : it->second();
  c->_class_name = class_name;
  return c;
}

std::vector<std::string> ClassRegistry::classNames() const {
  std::vector<std::string> v;
  for (const auto& kv : _factories) v.push_back(kv.first);
  return v;
}

namespace {

struct Parser {
  const std::string& s;
  const std::string& origin;
  size_t i = 0;
  int line = 1;

  [[noreturn]] void fail(const std::string& what) const {
    throw std::runtime_error("ConfigurableManager::read|" + origin + ":" + std::to_string(line) + ": " + what);
  }
  void skip() {
    for (;;) {
      while (i < s.size() && (s[i] == ' ' || s[i] == '\t' || s[i] == '\r' || s[i] == '\n')) {
        if (s[i] == '\n') ++line;
        ++i;
      }
      if (i + 1 < s.size() && s[i] == '/' && s[i + 1] == '/') {
        while (i < s.size() && s[i] != '\n') ++i;
        continue;
      }
      return;
    }
  }
  bool eof() {
    skip();
    return i >= s.size();
  }
  char peek() {
    skip();
    return i < s.size() ? s[i] : '\0';
  }
  void expect(char c) {
    if (peek() != c) fail(std::string("expected '") + c + "'");
    ++i;
  }
  std::string string() {
    if (peek() != '"') fail("expected a string");
    ++i;
    std::string out;
    while (i < s.size() && s[i] != '"') {
      if (s[i] == '\\' && i + 1 < s.size()) {
        ++i;
        switch (s[i]) {
          case 'n': out += '\n'; break;
          case 't': out += '\t'; break;
          default: out += s[i];
        }
      } else {
        if (s[i] == '\n') ++line;
        out += s[i];
      }
      ++i;
    }
    if (i >= s.size()) fail("unterminated string");
    ++i;
    return out;
  }
  ConfValue value() {
    const char c = peek();
    ConfValue v;
    if (c == '"') {
      v.kind = ConfValue::String;
      v.text = string();
    } else if (c == '[') {
      ++i;
      v.kind = ConfValue::Array;
      if (peek() == ']') {
        ++i;
        return v;
      }
      for (;;) {
        v.items.push_back(value());
        if (peek() == ',') {
          ++i;
          continue;
        }
        expect(']');
        break;
      }
    } else if (c == '{') {
      ++i;
      const std::string key = string();
      if (key != "#pointer") fail("only { \"#pointer\" : id } objects may be nested");
      expect(':');
      ConfValue id = value();
      if (id.kind != ConfValue::Number) fail("#pointer needs an integer");
      expect('}');
      v.kind = ConfValue::Pointer;
      v.pointer = (int) id.number;
    } else {
      const size_t b = i;
      while (i < s.size() && (std::isdigit((unsigned char) s[i]) || s[i] == '-' || s[i] == '+' || s[i] == '.' ||
                              s[i] == 'e' || s[i] == 'E' || s[i] == 'n' || s[i] == 'a' || s[i] == 'i' || s[i] == 'f'))
        ++i;
      if (b == i) fail("unexpected character");
      v.kind = ConfValue::Number;
      v.text = s.substr(b, i - b);
      char* end = nullptr;
      v.number = std::strtod(v.text.c_str(), &end);
      if (end == v.text.c_str()) fail("bad number '" + v.text + "'");
    }
    return v;
  }
};

struct RawObject {
  std::string class_name;
  int id = -1;
  std::vector<std::pair<std::string, ConfValue>> pairs;
  int line = 0;
};

std::string escape(const std::string& s) {
  std::string o;
  for (char c : s) {
    if (c == '"' || c == '\\') o += '\\';
    o += c;
  }
  return o;
}

void emit(std::ostream& os, const ConfValue& v, int indent) {
  switch (v.kind) {
    case ConfValue::Number: {
      if (!v.text.empty()) {
        os << v.text;
      } else if (v.number == std::floor(v.number) && std::fabs(v.number) < 1e15) {
        os << (long long) v.number;
      } else {
        char buf[64];
        std::snprintf(buf, sizeof buf, "%.9g", v.number);
        os << buf;
      }
      break;
    }
    case ConfValue::String: os << '"' << escape(v.text) << '"'; break;
    case ConfValue::Pointer:
      os << "{ \n" << std::string(indent + 2, ' ') << "\"#pointer\" : " << v.pointer << "\n" << std::string(indent + 1, ' ') << "}";
      break;
    case ConfValue::Array: {
      os << "[ ";
      for (size_t k = 0; k < v.items.size(); ++k) {
        if (k) os << ", ";
        emit(os, v.items[k], indent);
      }
      os << " ]";
      break;
    }
    case ConfValue::Null: os << "0"; break;
  }
}

}  // namespace

void ConfigurableManager::read(const std::string& filename) {
  std::ifstream f(filename);
  if (!f) throw std::runtime_error("ConfigurableManager::read|cannot open '" + filename + "'");
  std::stringstream ss;
  ss << f.rdbuf();
  readString(ss.str(), filename);
}

void ConfigurableManager::readString(const std::string& text, const std::string& origin) {
  Parser p{text, origin};
  std::vector<RawObject> raw;
  while (!p.eof()) {
    RawObject o;
    o.line = p.line;
    o.class_name = p.string();
    p.expect('{');
    if (p.peek() != '}') {
      for (;;) {
        const std::string key = p.string();
        p.expect(':');
        ConfValue v = p.value();
        if (key == "#id") {
          if (v.kind != ConfValue::Number) p.fail("#id needs an integer");
          o.id = (int) v.number;
        } else {
          o.pairs.emplace_back(key, std::move(v));
        }
        if (p.peek() == ',') {
          ++p.i;
          continue;
        }
        break;
      }
    }
    p.expect('}');
    if (o.id < 0) p.fail("object \"" + o.class_name + "\" has no #id");
    raw.push_back(std::move(o));
  }
  // pass 1: instantiate (forward references are legal), pass 2: assign values / resolve links
  std::map<int, ConfigurablePtr> by_id;
  std::vector<ConfigurablePtr> created;
  for (const RawObject& o : raw) {
    if (by_id.count(o.id)) throw std::runtime_error("ConfigurableManager::read|" + origin + ": duplicate #id " + std::to_string(o.id));
    ConfigurablePtr c = ClassRegistry::instance().create(o.class_name);
    by_id[o.id] = c;
    created.push_back(c);
  }
  Resolver resolve = [&](int id) -> ConfigurablePtr {
    auto it = by_id.find(id);
    if (it == by_id.end()) throw std::runtime_error("ConfigurableManager::read|" + origin + ": dangling #pointer " + std::to_string(id));
    return it->second;
  };
  for (size_t k = 0; k < raw.size(); ++k) {
    Configurable& c = *created[k];
    for (const auto& kv : raw[k].pairs) {
      if (kv.first == "name" && kv.second.kind == ConfValue::String) {
        c.setName(kv.second.text);
        continue;
      }
      if (PropertyBase* prop = c.property(kv.first)) {
        try {
          prop->fromConf(kv.second, resolve);
        } catch (const std::exception& e) {
          throw std::runtime_error("ConfigurableManager::read|" + origin + ": \"" + raw[k].class_name + "\" #id " +
                                   std::to_string(raw[k].id) + ": " + e.what());
        }
      } else {
        c.extraValues()[kv.first] = kv.second;
      }
    }
    _instances.push_back(created[k]);
    _ids.push_back(raw[k].id);
  }
}

ConfigurablePtr ConfigurableManager::create(const std::string& class_name, const std::string& name) {
  ConfigurablePtr c = ClassRegistry::instance().create(class_name);
  c->setName(name);
  int id = 1;
  for (int i : _ids) id = std::max(id, i + 1);
  _instances.push_back(c);
  _ids.push_back(id);
  return c;
}

ConfigurablePtr ConfigurableManager::getByName(const std::string& name) const {
  for (const ConfigurablePtr& c : _instances)
    if (c->name() == name) return c;
  return nullptr;
}

ConfigurablePtr ConfigurableManager::getById(int id) const {
  for (size_t k = 0; k < _instances.size(); ++k)
    if (_ids[k] == id) return _instances[k];
  return nullptr;
}

int ConfigurableManager::idOf(const Configurable* c) const {
  for (size_t k = 0; k < _instances.size(); ++k)
    if (_instances[k].get() == c) return _ids[k];
  return -1;
}

std::string ConfigurableManager::writeString(const std::vector<ConfigurablePtr>& roots) const {
  // closure of the roots over their links, in instance order
  std::set<const Configurable*> keep;
  if (!roots.empty()) {
    std::vector<ConfigurablePtr> todo = roots;
    while (!todo.empty()) {
      ConfigurablePtr c = todo.back();
      todo.pop_back();
      if (!c || keep.count(c.get())) continue;
      keep.insert(c.get());
      for (PropertyBase* p : c->properties()) p->collectLinks(todo);
      for (const auto& kv : c->extraValues()) {
        std::vector<const ConfValue*> vs{&kv.second};
        while (!vs.empty()) {
          const ConfValue* v = vs.back();
          vs.pop_back();
          if (v->kind == ConfValue::Pointer && v->pointer >= 0) todo.push_back(getById(v->pointer));
          for (const ConfValue& i : v->items) vs.push_back(&i);
        }
      }
    }
  }
  IdOf id_of = [this](const Configurable* c) {
    const int id = idOf(c);
    if (id < 0) throw std::runtime_error("ConfigurableManager::write|linked object is not owned by this manager");
    return id;
  };
  std::ostringstream os;
  for (size_t k = 0; k < _instances.size(); ++k) {
    const Configurable& c = *_instances[k];
    if (!roots.empty() && !keep.count(&c)) continue;
    os << '"' << c.className() << "\" { \n  \"#id\" : " << _ids[k];
    if (!c.name().empty()) os << ", \n  \"name\" : \"" << escape(c.name()) << '"';
    // declared PARAMs and pass-through keys, sorted by key like the reference's files
    std::map<std::string, std::pair<std::string, ConfValue>> out;
    for (PropertyBase* p : c.properties()) out[p->name()] = {p->description(), p->toConf(id_of)};
    for (const auto& kv : c.extraValues())
      if (!out.count(kv.first)) out[kv.first] = {"", kv.second};
    for (const auto& kv : out) {
      os << ", \n";
      if (!kv.second.first.empty()) os << "\n  // " << kv.second.first << "\n";
      os << "  \"" << kv.first << "\" : ";
      emit(os, kv.second.second, 2);
    }
    os << "\n }\n\n";
  }
  return os.str();
}

void ConfigurableManager::write(const std::string& filename, const std::vector<ConfigurablePtr>& roots) const {
  std::ofstream f(filename);
  if (!f) throw std::runtime_error("ConfigurableManager::write|cannot open '" + filename + "'");
  f << writeString(roots);
}

}  // namespace pslam_host
