// cli_options.h -- minimal command-line parser for the two CLIs.  Accepts the option spellings the
// reference's cxxopts declarations accept (associate_kmers.cpp:38-53, emma_kinship_kmers.cpp:37-42):
// "--name value", "--name=value", "-n value", and value-less boolean flags.
#ifndef KGH_CLI_OPTIONS_H
#define KGH_CLI_OPTIONS_H

#include <map>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

class CliOptions {
	public:
		struct ParseError : std::runtime_error {
			explicit ParseError(const std::string &m) : std::runtime_error(m) {}
		};
		CliOptions(const std::string &prog, const std::string &desc) : m_prog(prog), m_desc(desc) {}
		// short_name may be 0.  is_flag: takes no value.  def: default value ("" = none).
		void add(char short_name, const std::string &name, const std::string &help, bool is_flag = false,
		         const std::string &def = "") {
			Opt o;
			o.short_name = short_name; o.name = name; o.help = help; o.is_flag = is_flag; o.def = def;
			m_opts.push_back(o);
		}
		void parse(int argc, char **argv) {
			for (int i = 1; i < argc; i++) {
				std::string a = argv[i];
				const Opt *o = nullptr;
				std::string val;
				bool has_val = false;
				if (a.size() > 2 && a[0] == '-' && a[1] == '-') {
					const size_t eq = a.find('=');
					const std::string name = a.substr(2, eq == std::string::npos ? std::string::npos : eq - 2);
					o = find_long(name);
					if (!o) throw ParseError("Option '" + name + "' does not exist");
					if (eq != std::string::npos) { val = a.substr(eq + 1); has_val = true; }
				} else if (a.size() >= 2 && a[0] == '-' && a[1] != '-') {
					o = find_short(a[1]);
					if (!o) throw ParseError(std::string("Option '") + a[1] + "' does not exist");
					if (a.size() > 2) { val = a.substr(2); has_val = true; }
				} else {
					throw ParseError("Unexpected argument '" + a + "'");
				}
				if (o->is_flag) {
					m_vals[o->name] = "true";
					continue;
				}
				if (!has_val) {
					if (i + 1 >= argc) throw ParseError("Option '" + o->name + "' is missing an argument");
					val = argv[++i];
				}
				m_vals[o->name] = val;
			}
		}
		size_t count(const std::string &name) const { return m_vals.count(name); }
		bool has_value(const std::string &name) const {
			if (m_vals.count(name)) return true;
			const Opt *o = find_long(name);
			return o && !o->def.empty();
		}
		std::string str(const std::string &name) const {
			auto it = m_vals.find(name);
			if (it != m_vals.end()) return it->second;
			const Opt *o = find_long(name);
			if (o && !o->def.empty()) return o->def;
			throw ParseError("Option '" + name + "' has no value");
		}
		template <class T>
		T as(const std::string &name) const {
			std::istringstream is(str(name));
			T v;
			if (!(is >> v) || !is.eof()) throw ParseError("Argument '" + str(name) + "' failed to parse for option '" + name + "'");
			return v;
		}
		std::string help() const {
			std::ostringstream os;
			os << m_desc << "\nUsage:\n  " << m_prog << " [OPTION...]\n\n";
			for (const Opt &o : m_opts) {
				std::string left = "  ";
				left += o.short_name ? std::string("-") + o.short_name + ", " : std::string("    ");
				left += "--" + o.name + (o.is_flag ? "" : " arg");
				if (left.size() < 30) left.resize(30, ' ');
				os << left << " " << o.help;
				if (!o.def.empty()) os << " (default: " << o.def << ")";
				os << "\n";
			}
			return os.str();
		}
	private:
		struct Opt { char short_name; std::string name, help, def; bool is_flag; };
		const Opt *find_long(const std::string &n) const {
			for (const Opt &o : m_opts) if (o.name == n) return &o;
			return nullptr;
		}
		const Opt *find_short(char c) const {
			for (const Opt &o : m_opts) if (o.short_name == c) return &o;
			return nullptr;
		}
		std::string m_prog, m_desc;
		std::vector<Opt> m_opts;
		std::map<std::string, std::string> m_vals;
};

template <>
inline std::string CliOptions::as<std::string>(const std::string &name) const { return str(name); }

#endif
