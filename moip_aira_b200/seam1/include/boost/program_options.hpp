// boost/program_options.hpp -- header-only stand-in for the slice of Boost.Program_options that moip_aira's
// main() uses (reference src/aira.cpp:138-215).  NOT Boost: written from the reference's call
// sites so that the UNMODIFIED src/aira.cpp builds in an image without Boost headers.  Supported surface:
//   options_description(caption), add_options()(name, description) / (name, value_semantic*, description),
//   value<T>(T*), bool_switch(bool*), ->default_value(v), parse_command_line, store, notify,
//   variables_map::count(name), operator<<(ostream, options_description).
// Command-line syntax accepted (what Boost's default style accepts for these options): --long value, --long=value,
// unambiguous --long prefixes, -s value, -svalue, bool switches without a value; an unknown option or a missing
// value throws boost::program_options::error (uncaught in the reference, i.e. the program terminates).
#ifndef MOIP_B200_SEAM1_PROGRAM_OPTIONS_HPP
#define MOIP_B200_SEAM1_PROGRAM_OPTIONS_HPP

#include <map>
#include <memory>
#include <ostream>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

namespace boost {
namespace program_options {

class error : public std::logic_error {
 public:
  explicit error(const std::string& what) : std::logic_error(what) {}
};

class value_semantic {
 public:
  virtual ~value_semantic() {}
  virtual bool takes_value() const = 0;
  virtual bool has_default() const = 0;
  virtual void apply_default() = 0;
  virtual void parse(const std::string& option, const std::string& token) = 0;
  virtual void set_switch() = 0;
  virtual std::string default_text() const = 0;
};

template <class T>
class typed_value : public value_semantic {
 public:
  explicit typed_value(T* store, bool is_switch = false) : store_(store), switch_(is_switch) {}
  typed_value* default_value(const T& v) {
    default_ = v;
    has_default_ = true;
    return this;
  }
  bool takes_value() const override { return !switch_; }
  bool has_default() const override { return has_default_; }
  void apply_default() override {
    if (store_ && has_default_) *store_ = default_;
  }
  void parse(const std::string& option, const std::string& token) override {
    T v{};
    if (!convert(token, v)) throw error("the argument ('" + token + "') for option '--" + option + "' is invalid");
    if (store_) *store_ = v;
  }
  void set_switch() override { assign_true(store_); }
  std::string default_text() const override {
    if (!has_default_ || switch_) return "";
    std::ostringstream ss;
    ss << " (=" << default_ << ")";
    return ss.str();
  }

 private:
  template <class U>
  static bool convert(const std::string& token, U& v) {
    std::istringstream ss(token);
    ss >> v;
    return !ss.fail() && (ss >> std::ws).eof();
  }
  static bool convert(const std::string& token, std::string& v) {
    v = token;
    return true;
  }
  static void assign_true(bool* p) {
    if (p) *p = true;
  }
  template <class U>
  static void assign_true(U*) {}
  T* store_;
  bool switch_;
  bool has_default_ = false;
  T default_{};
};

template <class T>
typed_value<T>* value(T* store) {
  return new typed_value<T>(store);
}
inline typed_value<bool>* bool_switch(bool* store) {
  return (new typed_value<bool>(store, true))->default_value(false);
}

struct option_description {
  std::string long_name;
  char short_name = 0;
  std::string description;
  std::shared_ptr<value_semantic> semantic;  // null: a plain flag such as --help
};

class options_description;

class options_description_easy_init {
 public:
  explicit options_description_easy_init(options_description* owner) : owner_(owner) {}
  options_description_easy_init& operator()(const char* name, const char* description);
  options_description_easy_init& operator()(const char* name, value_semantic* s, const char* description);

 private:
  options_description* owner_;
};

class options_description {
 public:
  explicit options_description(const std::string& caption) : caption_(caption) {}
  options_description_easy_init add_options() { return options_description_easy_init(this); }
  void add(const char* name, value_semantic* s, const char* description) {
    option_description d;
    std::string spec(name);
    const size_t comma = spec.find(',');
    d.long_name = spec.substr(0, comma);
    if (comma != std::string::npos && comma + 1 < spec.size()) d.short_name = spec[comma + 1];
    d.description = description;
    d.semantic.reset(s);
    options_.push_back(d);
  }
  const std::vector<option_description>& options() const { return options_; }
  const std::string& caption() const { return caption_; }
  const option_description* find_long(const std::string& name) const {
    const option_description* prefix_hit = nullptr;
    int prefix_hits = 0;
    for (const auto& o : options_) {
      if (o.long_name == name) return &o;
      if (!name.empty() && o.long_name.compare(0, name.size(), name) == 0) {
        prefix_hit = &o;
        ++prefix_hits;
      }
    }
    if (prefix_hits > 1) throw error("option '--" + name + "' is ambiguous");
    return prefix_hit;
  }
  const option_description* find_short(char c) const {
    for (const auto& o : options_)
      if (o.short_name == c) return &o;
    return nullptr;
  }

 private:
  std::string caption_;
  std::vector<option_description> options_;
};

inline options_description_easy_init& options_description_easy_init::operator()(const char* name,
                                                                               const char* description) {
  owner_->add(name, nullptr, description);
  return *this;
}
inline options_description_easy_init& options_description_easy_init::operator()(const char* name, value_semantic* s,
                                                                               const char* description) {
  owner_->add(name, s, description);
  return *this;
}

inline std::ostream& operator<<(std::ostream& os, const options_description& d) {
  os << d.caption() << ":\n";
  for (const auto& o : d.options()) {
    std::string head = "  ";
    if (o.short_name) head += std::string("-") + o.short_name + " [ --" + o.long_name + " ]";
    else head += "--" + o.long_name;
    if (o.semantic && o.semantic->takes_value()) head += " arg" + o.semantic->default_text();
    if (head.size() < 28) head.resize(28, ' ');
    else head += ' ';
    os << head;
    // continuation lines of a description are indented under the first one
    std::string desc = o.description;
    size_t pos = 0;
    while ((pos = desc.find('\n', pos)) != std::string::npos) {
      desc.insert(pos + 1, std::string(28, ' '));
      pos += 29;
    }
    os << desc << "\n";
  }
  return os;
}

// one recognised option occurrence: its description and the value token (empty for flags/switches)
struct parsed_option {
  const option_description* desc;
  std::string value;
};
struct parsed_options {
  const options_description* description;
  std::vector<parsed_option> options;
};

inline parsed_options parse_command_line(int argc, const char* const* argv, const options_description& desc) {
  parsed_options out;
  out.description = &desc;
  for (int i = 1; i < argc; ++i) {
    const std::string tok(argv[i]);
    const option_description* od = nullptr;
    std::string val;
    bool have_val = false;
    if (tok.size() > 2 && tok[0] == '-' && tok[1] == '-') {
      const size_t eq = tok.find('=');
      const std::string name = tok.substr(2, eq == std::string::npos ? std::string::npos : eq - 2);
      od = desc.find_long(name);
      if (!od) throw error("unrecognised option '" + tok + "'");
      if (eq != std::string::npos) {
        val = tok.substr(eq + 1);
        have_val = true;
      }
    } else if (tok.size() >= 2 && tok[0] == '-' && tok[1] != '-') {
      od = desc.find_short(tok[1]);
      if (!od) throw error("unrecognised option '" + tok + "'");
      if (tok.size() > 2) {
        val = tok.substr(2);
        have_val = true;
      }
    } else {
      throw error("too many positional options have been specified on the command line");
    }
    const bool wants = od->semantic && od->semantic->takes_value();
    if (wants && !have_val) {
      if (i + 1 >= argc) throw error("the required argument for option '--" + od->long_name + "' is missing");
      val = argv[++i];
    } else if (!wants && have_val) {
      throw error("option '--" + od->long_name + "' does not take any arguments");
    }
    out.options.push_back(parsed_option{od, val});
  }
  return out;
}
inline parsed_options parse_command_line(int argc, char** argv, const options_description& desc) {
  return parse_command_line(argc, const_cast<const char* const*>(argv), desc);
}

class variables_map {
 public:
  size_t count(const std::string& name) const {
    auto it = seen_.find(name);
    return it == seen_.end() ? 0 : 1;
  }
  // occurrences given on the command line win; defaults fill in the rest (and count as present, like Boost)
  void absorb(const parsed_options& p) {
    for (const auto& o : p.options) {
      if (given_.count(o.desc->long_name)) throw error("option '--" + o.desc->long_name + "' cannot be specified more than once");
      given_[o.desc->long_name] = 1;
      seen_[o.desc->long_name] = 1;
      if (!o.desc->semantic) continue;
      if (o.desc->semantic->takes_value()) o.desc->semantic->parse(o.desc->long_name, o.value);
      else pending_switch_.push_back(o.desc);
    }
    for (const auto& d : p.description->options()) {
      if (!d.semantic || given_.count(d.long_name)) continue;
      if (d.semantic->has_default()) {
        seen_[d.long_name] = 1;
        pending_default_.push_back(&d);
      }
    }
  }
  void finish() {
    for (const option_description* d : pending_default_) d->semantic->apply_default();
    for (const option_description* d : pending_switch_) d->semantic->set_switch();
    pending_default_.clear();
    pending_switch_.clear();
  }

 private:
  std::map<std::string, int> seen_, given_;
  std::vector<const option_description*> pending_default_, pending_switch_;
};

inline void store(const parsed_options& p, variables_map& vm) { vm.absorb(p); }
// Boost writes the bound variables (value<T>(&x)) in notify(); the reference reads them only afterwards
inline void notify(variables_map& vm) { vm.finish(); }

}  // namespace program_options
}  // namespace boost
#endif  // MOIP_B200_SEAM1_PROGRAM_OPTIONS_HPP
