// Minimal stand-in for <boost/program_options.hpp>: enough surface for the reference's command
// classes to COMPILE (variables_map lookups, option declarations).  No command line is ever parsed
// through it -- the oracle/ref driver constructs the command objects directly, like the
// reference's own tests do (src/testGossCmdBuildGraph.cc:120-137).
#pragma once
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <typeinfo>
#include <vector>
#include <boost/shared_ptr.hpp>
namespace boost { namespace program_options {
class variable_value {
public:
    variable_value() : defaulted_(false) {}
    template <typename T> explicit variable_value(const T& v, bool defaulted = false) : p_(std::make_shared<T>(v)), defaulted_(defaulted) {}
    bool empty() const { return !p_; }
    bool defaulted() const { return defaulted_; }
    template <typename T> const T& as() const { if (!p_) throw std::runtime_error("empty option value"); return *static_cast<const T*>(p_.get()); }
    template <typename T> T& as() { if (!p_) throw std::runtime_error("empty option value"); return *static_cast<T*>(p_.get()); }
private:
    std::shared_ptr<void> p_;
    bool defaulted_;
};
class variables_map : public std::map<std::string, variable_value> {
public:
    const variable_value& operator[](const std::string& name) const {
        static const variable_value empty;
        const_iterator it = find(name);
        return it == end() ? empty : it->second;
    }
    variable_value& operator[](const std::string& name) { return std::map<std::string, variable_value>::operator[](name); }
    void notify() {}
};
class value_semantic { public: virtual ~value_semantic() {} };
template <typename T> class typed_value : public value_semantic {
public:
    typed_value* default_value(const T&) { return this; }
    typed_value* default_value(const T&, const std::string&) { return this; }
    typed_value* implicit_value(const T&) { return this; }
    typed_value* multitoken() { return this; }
    typed_value* composing() { return this; }
    typed_value* zero_tokens() { return this; }
    typed_value* required() { return this; }
};
template <typename T> typed_value<T>* value() { return new typed_value<T>(); }
template <typename T> typed_value<T>* value(T*) { return new typed_value<T>(); }
inline typed_value<bool>* bool_switch() { return new typed_value<bool>(); }
inline typed_value<bool>* bool_switch(bool*) { return new typed_value<bool>(); }
class options_description;
class options_description_easy_init {
public:
    options_description_easy_init& operator()(const char*, const char*) { return *this; }
    options_description_easy_init& operator()(const char*, const value_semantic* s) { delete s; return *this; }
    options_description_easy_init& operator()(const char*, const value_semantic* s, const char*) { delete s; return *this; }
};
class options_description {
public:
    options_description() {}
    explicit options_description(const std::string&) {}
    options_description(const std::string&, unsigned) {}
    options_description_easy_init add_options() { return options_description_easy_init(); }
    options_description& add(const options_description&) { return *this; }
    friend std::ostream& operator<<(std::ostream& os, const options_description&) { return os; }
};
class positional_options_description {
public:
    positional_options_description& add(const char*, int) { return *this; }
};
struct parsed_options { };
class command_line_parser {
public:
    command_line_parser(int, const char* const*) {}
    command_line_parser(const std::vector<std::string>&) {}
    command_line_parser& options(const options_description&) { return *this; }
    command_line_parser& positional(const positional_options_description&) { return *this; }
    command_line_parser& allow_unregistered() { return *this; }
    parsed_options run() { return parsed_options(); }
};
inline void store(const parsed_options&, variables_map&) {}
inline void notify(variables_map&) {}
inline std::vector<std::string> collect_unrecognized(const parsed_options&, int) { return std::vector<std::string>(); }
enum collect_unrecognized_mode { include_positional, exclude_positional };
class error : public std::logic_error { public: explicit error(const std::string& w) : std::logic_error(w) {} };
}}  // namespace boost::program_options
#include <ostream>
