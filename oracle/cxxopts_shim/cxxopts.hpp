// Minimal stand-in for the part of cxxopts v3 the reference's src/main.cpp uses -- TEST INFRASTRUCTURE, written for
// this repository (cxxopts is a FetchContent download in the reference's CMakeLists.txt:109-116 and there is no
// network here).  Supports: Options(name[, help]); add_options()(name, description, value<T>()) chains;
// parse(argc, argv) with "--key=value", "--key value" and bare "--flag" (bool -> true); result.count(key);
// result[key].as<T>() for std::string, int, unsigned, double, bool.  Unknown options and missing values throw, like
// cxxopts does.  For dataset_converter/main_preprocess.cpp also: value<T>()->default_value("..."), "s,long" short
// options, std::vector<std::string> values, parse_positional({...}), set_width / set_tab_expansion (formatting no-ops).
#pragma once

#include <map>
#include <memory>
#include <sstream>
#include <stdexcept>
#include <string>
#include <type_traits>
#include <vector>

namespace cxxopts {

namespace exceptions {
struct exception : std::runtime_error { using std::runtime_error::runtime_error; };
struct no_such_option : exception { using exception::exception; };
struct missing_argument : exception { using exception::exception; };
struct option_has_no_value : exception { using exception::exception; };
struct incorrect_argument_type : exception { using exception::exception; };
}  // namespace exceptions

struct Value : std::enable_shared_from_this<Value> {
    bool is_bool = false;
    bool is_list = false;
    bool has_default = false;
    std::string default_text;
    std::shared_ptr<Value> default_value(const std::string &text) {
        has_default = true;
        default_text = text;
        return shared_from_this();
    }
};

template<class T>
struct is_string_vector : std::false_type {};
template<>
struct is_string_vector<std::vector<std::string>> : std::true_type {};

template<class T>
std::shared_ptr<Value> value() {
    auto v = std::make_shared<Value>();
    v->is_bool = std::is_same_v<T, bool>;
    v->is_list = is_string_vector<T>::value;
    return v;
}

class OptionValue {
    std::string key, text;
    std::vector<std::string> items;
    bool present = false, defaulted = false;
    friend class Options;
public:
    OptionValue() = default;
    OptionValue(std::string k, std::string t, bool from_default = false)
            : key(std::move(k)), text(std::move(t)), present(true), defaulted(from_default) { items.push_back(text); }
    std::size_t count() const { return present && !defaulted ? items.size() : 0; }

    template<class T>
    T as() const {
        if (!present) throw exceptions::option_has_no_value("Option '" + key + "' has no value");
        if constexpr (is_string_vector<T>::value) {
            return items;
        } else
        if constexpr (std::is_same_v<T, std::string>) {
            return text;
        } else if constexpr (std::is_same_v<T, bool>) {
            if (text == "true" || text == "True" || text == "1" || text == "t" || text == "T") return true;
            if (text == "false" || text == "False" || text == "0" || text == "f" || text == "F") return false;
            throw exceptions::incorrect_argument_type("Argument '" + text + "' failed to parse");
        } else {
            std::istringstream in(text);
            T v{};
            in >> v;
            if (in.fail() || !in.eof()) throw exceptions::incorrect_argument_type("Argument '" + text + "' failed to parse");
            return v;
        }
    }
};

class ParseResult {
    std::map<std::string, OptionValue> values;
    friend class Options;
public:
    ParseResult() = default;
    std::size_t count(const std::string &key) const {
        auto it = values.find(key);
        return it == values.end() ? 0 : it->second.count();
    }
    const OptionValue &operator[](const std::string &key) const {
        static const OptionValue none;
        auto it = values.find(key);
        if (it == values.end()) {
            // cxxopts throws when an option that was never given (and has no default) is read
            throw exceptions::option_has_no_value("Option '" + key + "' has no value");
        }
        return it->second;
    }
};

class Options;

class OptionAdder {
    Options &owner;
public:
    explicit OptionAdder(Options &o) : owner(o) {}
    OptionAdder &operator()(const std::string &names, const std::string &description,
                            const std::shared_ptr<Value> &v = value<bool>());
};

class Options {
    std::string program, description;
    struct Spec { std::string description; bool is_bool; bool is_list; bool has_default; std::string default_text; };
    std::map<std::string, Spec> specs;
    std::map<std::string, std::string> short_to_long;
    std::vector<std::string> positional;
    friend class OptionAdder;
public:
    explicit Options(std::string name, std::string help = "") : program(std::move(name)), description(std::move(help)) {}

    Options &set_width(std::size_t) { return *this; }
    Options &set_tab_expansion(bool = true) { return *this; }
    void parse_positional(std::initializer_list<std::string> names) { positional.assign(names.begin(), names.end()); }

    OptionAdder add_options(const std::string & = "") { return OptionAdder(*this); }

    ParseResult parse(int argc, const char *const *argv) const {
        ParseResult r;
        for (const auto &kv: specs)
            if (kv.second.has_default) r.values[kv.first] = OptionValue(kv.first, kv.second.default_text, true);
        auto add = [&r, this](const std::string &key, const std::string &val) {
            auto it = r.values.find(key);
            if (specs.at(key).is_list && it != r.values.end() && !it->second.defaulted) it->second.items.push_back(val);
            else r.values[key] = OptionValue(key, val);
        };
        std::size_t next_positional = 0;
        for (int i = 1; i < argc; ++i) {
            std::string arg = argv[i];
            if (arg.rfind("--", 0) != 0) {
                if (arg.size() == 2 && arg[0] == '-' && short_to_long.count(arg.substr(1))) {
                    arg = "--" + short_to_long.at(arg.substr(1));
                } else {
                    // positional: goes to the next name given to parse_positional (a list option takes all the rest)
                    if (next_positional < positional.size()) {
                        const std::string &key = positional[next_positional];
                        add(key, arg);
                        if (!specs.at(key).is_list) ++next_positional;
                    }
                    continue;
                }
            }
            std::string key = arg.substr(2), val;
            bool has_val = false;
            auto eq = key.find('=');
            if (eq != std::string::npos) {
                val = key.substr(eq + 1);
                key = key.substr(0, eq);
                has_val = true;
            }
            auto it = specs.find(key);
            if (it == specs.end()) throw exceptions::no_such_option("Option '" + key + "' does not exist");
            if (!has_val) {
                if (it->second.is_bool) {
                    val = "true";
                } else {
                    if (i + 1 >= argc) throw exceptions::missing_argument("Option '" + key + "' is missing an argument");
                    val = argv[++i];
                }
            }
            add(key, val);
        }
        return r;
    }

    std::string help() const {
        std::string s = description + "\nUsage:\n  " + program + " [OPTION...]\n\n";
        for (const auto &kv: specs) s += "      --" + kv.first + " arg  " + kv.second.description + "\n";
        return s;
    }
};

inline OptionAdder &OptionAdder::operator()(const std::string &names, const std::string &description,
                                            const std::shared_ptr<Value> &v) {
    std::string longname = names;
    auto comma = names.find(',');
    if (comma != std::string::npos) {
        std::string a = names.substr(0, comma), b = names.substr(comma + 1);
        longname = a.size() > b.size() ? a : b;
        owner.short_to_long[a.size() > b.size() ? b : a] = longname;
    }
    owner.specs[longname] = Options::Spec{description, v->is_bool, v->is_list, v->has_default, v->default_text};
    return *this;
}

}  // namespace cxxopts
