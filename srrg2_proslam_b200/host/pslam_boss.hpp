// pslam_boss.hpp -- minimal host-side mirror of the srrg2_core plugin machinery the frontend path uses:
// Configurable + PARAM properties + class registry + the BOSS `.conf` reader / writer.
//
// The reference selects every module by the class-name string in its `.conf`
// (srrg2_proslam/src/srrg2_proslam/sensor_processing/instances.cpp:11-17,
//  srrg2_proslam/src/srrg2_proslam/registration/instances.cpp:51-76) and links modules with
// `"#pointer" : id` (configurations/kitti.conf:51-75).  srrg2_core itself is not in this image, so this file
// provides just enough of that interface (same PARAM names, same file grammar -- SURVEY.md App. C) for the
// CUDA-backed classes of pslam_plugin.hpp to be instantiated from the UNCHANGED kitti.conf / icl.conf /
// euroc.conf.  Classes that are not on the hot path are loaded as GenericConfigurable (all values kept, so a
// file round-trips) -- they are control plane and out of scope.
#pragma once
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <type_traits>
#include <functional>
#include <map>
#include <memory>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

namespace pslam_host {

// ---- a parsed value of the .conf grammar --------------------------------------------------------------------
struct ConfValue {
  enum Kind { Null, Number, String, Array, Pointer } kind = Null;
  double number = 0;
  std::string text;  // String: the string; Number: the literal as written (kept for exact round trips)
  std::vector<ConfValue> items;
  int pointer = -1;
  static ConfValue num(double v) {
    ConfValue c;
    c.kind = Number;
    c.number = v;
    return c;
  }
  static ConfValue str(const std::string& s) {
    ConfValue c;
    c.kind = String;
    c.text = s;
    return c;
  }
  static ConfValue ptr(int id) {
    ConfValue c;
    c.kind = Pointer;
    c.pointer = id;
    return c;
  }
};

class Configurable;
class PropertyBase;
using ConfigurablePtr = std::shared_ptr<Configurable>;
using Resolver = std::function<ConfigurablePtr(int)>;
using IdOf = std::function<int(const Configurable*)>;

// ---- Configurable -------------------------------------------------------------------------------------------
class Configurable : public std::enable_shared_from_this<Configurable> {
public:
  virtual ~Configurable() = default;
  const std::string& className() const { return _class_name; }
  const std::string& name() const { return _name; }
  void setName(const std::string& n) { _name = n; }
  PropertyBase* property(const std::string& name) const;
  const std::vector<PropertyBase*>& properties() const { return _properties; }
  // keys present in a file but not declared as PARAM (base-class params of modules we only partly mirror)
  std::map<std::string, ConfValue>& extraValues() { return _extra; }
  const std::map<std::string, ConfValue>& extraValues() const { return _extra; }

private:
  friend class PropertyBase;
  friend class ConfigurableManager;
  friend class ClassRegistry;
  std::string _class_name, _name;
  std::vector<PropertyBase*> _properties;
  std::map<std::string, ConfValue> _extra;
};

// a module outside the hot path: keeps every key / value of its .conf block
class GenericConfigurable : public Configurable {};


// ---- properties (srrg2_core::Property*) ---------------------------------------------------------------------
class PropertyBase {
public:
  PropertyBase(const char* name, const char* description, Configurable* owner, bool* changed_flag);
  virtual ~PropertyBase() = default;
  const std::string& name() const { return _name; }
  const std::string& description() const { return _description; }
  virtual void fromConf(const ConfValue& v, const Resolver& resolve) = 0;
  virtual ConfValue toConf(const IdOf& id_of) const = 0;
  virtual void collectLinks(std::vector<ConfigurablePtr>&) const {}

protected:
  void touch() {
    if (_changed_flag) *_changed_flag = true;
  }
  std::string _name, _description;
  bool* _changed_flag;
};

template <typename T>
class PropertyScalar_ : public PropertyBase {
public:
  PropertyScalar_(const char* n, const char* d, Configurable* o, const T& def, bool* flag = nullptr)
    : PropertyBase(n, d, o, flag), _value(def) {}
  const T& value() const { return _value; }
  void setValue(const T& v) {
    _value = v;
    touch();
  }
  void fromConf(const ConfValue& v, const Resolver&) override {
    if (v.kind != ConfValue::Number) throw std::runtime_error("property '" + _name + "': number expected");
    _value = static_cast<T>(v.number);
    touch();
  }
  ConfValue toConf(const IdOf&) const override {
    ConfValue c = ConfValue::num(static_cast<double>(_value));
    if (std::is_same<T, float>::value) {  // shortest literal that reads back as the same float (0.1f -> "0.1")
      char buf[64];
      for (int prec = 1; prec <= 9; ++prec) {
        std::snprintf(buf, sizeof buf, "%.*g", prec, static_cast<double>(_value));
        if (static_cast<T>(std::strtod(buf, nullptr)) == _value) break;
      }
      c.text = buf;
    }
    return c;
  }

protected:
  T _value;
};
using PropertyFloat = PropertyScalar_<float>;
using PropertyDouble = PropertyScalar_<double>;
using PropertyInt = PropertyScalar_<int>;
using PropertyUnsignedInt = PropertyScalar_<unsigned>;
using PropertyBool = PropertyScalar_<bool>;

class PropertyString : public PropertyBase {
public:
  PropertyString(const char* n, const char* d, Configurable* o, const std::string& def, bool* flag = nullptr)
    : PropertyBase(n, d, o, flag), _value(def) {}
  const std::string& value() const { return _value; }
  void setValue(const std::string& v) {
    _value = v;
    touch();
  }
  void fromConf(const ConfValue& v, const Resolver&) override {
    if (v.kind != ConfValue::String) throw std::runtime_error("property '" + _name + "': string expected");
    _value = v.text;
    touch();
  }
  ConfValue toConf(const IdOf&) const override { return ConfValue::str(_value); }

private:
  std::string _value;
};

// PropertyVector_<T> / PropertyEigen_<VectorNf>: a flat array of numbers
template <typename T>
class PropertyVector_ : public PropertyBase {
public:
  PropertyVector_(const char* n, const char* d, Configurable* o, const std::vector<T>& def, bool* flag = nullptr)
    : PropertyBase(n, d, o, flag), _value(def) {}
  const std::vector<T>& value() const { return _value; }
  void setValue(const std::vector<T>& v) {
    _value = v;
    touch();
  }
  void pushBack(const T& v) {
    _value.push_back(v);
    touch();
  }
  void fromConf(const ConfValue& v, const Resolver&) override {
    if (v.kind != ConfValue::Array) throw std::runtime_error("property '" + _name + "': array expected");
    _value.clear();
    for (const ConfValue& i : v.items) {
      if (i.kind != ConfValue::Number) throw std::runtime_error("property '" + _name + "': array of numbers expected");
      _value.push_back(static_cast<T>(i.number));
    }
    touch();
  }
  ConfValue toConf(const IdOf&) const override {
    ConfValue c;
    c.kind = ConfValue::Array;
    for (const T& x : _value) c.items.push_back(ConfValue::num(static_cast<double>(x)));
    return c;
  }

private:
  std::vector<T> _value;
};

// PropertyConfigurable_<T>: `"#pointer" : id` link to another module (-1 = null)
class PropertyConfigurableBase : public PropertyBase {
public:
  using PropertyBase::PropertyBase;
  virtual void setPointer(const ConfigurablePtr& p) = 0;
  virtual ConfigurablePtr pointer() const = 0;
};
template <typename T>
class PropertyConfigurable_ : public PropertyConfigurableBase {
public:
  PropertyConfigurable_(const char* n, const char* d, Configurable* o, std::shared_ptr<T> def, bool* flag = nullptr)
    : PropertyConfigurableBase(n, d, o, flag), _value(std::move(def)) {}
  const std::shared_ptr<T>& value() const { return _value; }
  T* operator->() const { return _value.get(); }
  void setValue(std::shared_ptr<T> v) {
    _value = std::move(v);
    touch();
  }
  void setPointer(const ConfigurablePtr& p) override {
    if (!p) {
      _value.reset();
    } else {
      std::shared_ptr<T> t = std::dynamic_pointer_cast<T>(p);
      if (!t) throw std::runtime_error("property '" + _name + "': linked object has the wrong class");
      _value = t;
    }
    touch();
  }
  ConfigurablePtr pointer() const override { return std::static_pointer_cast<Configurable>(_value); }
  void fromConf(const ConfValue& v, const Resolver& resolve) override {
    if (v.kind != ConfValue::Pointer) throw std::runtime_error("property '" + _name + "': #pointer expected");
    _unmirrored.reset();
    ConfigurablePtr target = v.pointer < 0 ? nullptr : resolve(v.pointer);
    if (target && !std::dynamic_pointer_cast<T>(target) && std::dynamic_pointer_cast<GenericConfigurable>(target)) {
      // the file links a class this build does not mirror (e.g. the KD-tree finder): the link is kept for
      // round trips, the module that owns this property reports "not set" when it is used
      _unmirrored = target;
      _value.reset();
      touch();
      return;
    }
    setPointer(target);
  }
  ConfValue toConf(const IdOf& id_of) const override {
    if (_value) return ConfValue::ptr(id_of(_value.get()));
    return ConfValue::ptr(_unmirrored ? id_of(_unmirrored.get()) : -1);
  }
  void collectLinks(std::vector<ConfigurablePtr>& out) const override {
    if (_value) out.push_back(pointer());
    if (_unmirrored) out.push_back(_unmirrored);
  }
  const ConfigurablePtr& unmirrored() const { return _unmirrored; }

private:
  std::shared_ptr<T> _value;
  ConfigurablePtr _unmirrored;
};

template <typename T>
class PropertyConfigurableVector_ : public PropertyBase {
public:
  PropertyConfigurableVector_(const char* n, const char* d, Configurable* o, bool* flag = nullptr)
    : PropertyBase(n, d, o, flag) {}
  const std::vector<std::shared_ptr<T>>& value() const { return _value; }
  void pushBack(std::shared_ptr<T> v) {
    _value.push_back(std::move(v));
    touch();
  }
  size_t size() const { return _value.size(); }
  // srrg2_core PropertyConfigurableVector_::setValue(index, pointer) (tests/test_aligners.cpp:901)
  void setValue(size_t index, std::shared_ptr<T> v) {
    if (index >= _value.size()) throw std::runtime_error("property '" + _name + "': index out of range");
    _value[index] = std::move(v);
    touch();
  }
  void fromConf(const ConfValue& v, const Resolver& resolve) override {
    if (v.kind != ConfValue::Array) throw std::runtime_error("property '" + _name + "': array of #pointer expected");
    _value.clear();
    for (const ConfValue& i : v.items) {
      if (i.kind != ConfValue::Pointer) throw std::runtime_error("property '" + _name + "': array of #pointer expected");
      if (i.pointer < 0) {
        _value.push_back(nullptr);
        continue;
      }
      std::shared_ptr<T> t = std::dynamic_pointer_cast<T>(resolve(i.pointer));
      if (!t) throw std::runtime_error("property '" + _name + "': linked object has the wrong class");
      _value.push_back(t);
    }
    touch();
  }
  ConfValue toConf(const IdOf& id_of) const override {
    ConfValue c;
    c.kind = ConfValue::Array;
    for (const auto& p : _value) c.items.push_back(ConfValue::ptr(p ? id_of(p.get()) : -1));
    return c;
  }
  void collectLinks(std::vector<ConfigurablePtr>& out) const override {
    for (const auto& p : _value)
      if (p) out.push_back(std::static_pointer_cast<Configurable>(p));
  }

private:
  std::vector<std::shared_ptr<T>> _value;
};

// same spelling as srrg2_core's PARAM macro: PARAM(type, name, description, default, changed-flag pointer)
#define PARAM(TYPE, NAME, DESC, DEFAULT, FLAG) TYPE param_##NAME{#NAME, DESC, this, DEFAULT, FLAG}
#define PARAM_VECTOR(TYPE, NAME, DESC, FLAG) TYPE param_##NAME{#NAME, DESC, this, FLAG}

// ---- class registry (BOSS_REGISTER_CLASS) -------------------------------------------------------------------
class ClassRegistry {
public:
  using Factory = std::function<ConfigurablePtr()>;
  static ClassRegistry& instance();
  void add(const std::string& class_name, Factory f) { _factories[class_name] = std::move(f); }
  bool has(const std::string& class_name) const { return _factories.count(class_name) != 0; }
  ConfigurablePtr create(const std::string& class_name) const;  // GenericConfigurable when unknown
  std::vector<std::string> classNames() const;

private:
  std::map<std::string, Factory> _factories;
};
#define PSLAM_REGISTER_CLASS_AS(CLASS, NAME) \
  ::pslam_host::ClassRegistry::instance().add(NAME, [] { return std::static_pointer_cast<::pslam_host::Configurable>(std::make_shared<CLASS>()); })

// ---- ConfigurableManager: read / write a .conf, look modules up by name --------------------------------------
class ConfigurableManager {
public:
  void read(const std::string& filename);
  void readString(const std::string& text, const std::string& origin = "<string>");
  // writes `roots` and everything they link to (all instances when empty)
  void write(const std::string& filename, const std::vector<ConfigurablePtr>& roots = {}) const;
  std::string writeString(const std::vector<ConfigurablePtr>& roots = {}) const;
  ConfigurablePtr create(const std::string& class_name, const std::string& name = "");
  ConfigurablePtr getByName(const std::string& name) const;
  template <typename T>
  std::shared_ptr<T> getByName(const std::string& name) const {
    return std::dynamic_pointer_cast<T>(getByName(name));
  }
  ConfigurablePtr getById(int id) const;
  int idOf(const Configurable* c) const;
  const std::vector<ConfigurablePtr>& instances() const { return _instances; }

private:
  std::vector<ConfigurablePtr> _instances;
  std::vector<int> _ids;
};

}  // namespace pslam_host
