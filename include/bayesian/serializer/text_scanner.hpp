// bayesian/serializer/text_scanner.hpp — the character-level scanner shared by the BIF and DSC
// loaders (serializer/bif.hpp, serializer/dsc.hpp).  Hand-written: the reference's BIF grammar is a
// Boost.Spirit.Qi grammar (bayesian/serializer/bif.hpp:138-263) and its DSC reader works on whole
// lines with fixed column offsets (dsc.hpp:33-232); neither Boost nor fixed columns are wanted here.
#ifndef BNB200_BAYESIAN_SERIALIZER_TEXT_SCANNER_HPP
#define BNB200_BAYESIAN_SERIALIZER_TEXT_SCANNER_HPP

#include <cctype>
#include <cstddef>
#include <cstdlib>
#include <stdexcept>
#include <string>

namespace bn {
namespace serializer {

class text_scanner {
public:
    explicit text_scanner(std::string text, char const* format) : text_(std::move(text)), format_(format) {}

    // Skips white space (the reference's qi::ascii::space skipper) and, beyond the reference,
    // `// ...` and `/* ... */` comments.
    void skip()
    {
        for (;;) {
            while (pos_ < text_.size() && std::isspace(static_cast<unsigned char>(text_[pos_]))) ++pos_;
            if (pos_ + 1 < text_.size() && text_[pos_] == '/' && text_[pos_ + 1] == '/') {
                while (pos_ < text_.size() && text_[pos_] != '\n') ++pos_;
            } else if (pos_ + 1 < text_.size() && text_[pos_] == '/' && text_[pos_ + 1] == '*') {
                std::size_t const close = text_.find("*/", pos_ + 2);
                if (close == std::string::npos) fail("unterminated comment");
                pos_ = close + 2;
            } else {
                return;
            }
        }
    }

    bool at_end()
    {
        skip();
        return pos_ >= text_.size();
    }

    // the next non-blank character without consuming it (0 at the end)
    char peek()
    {
        skip();
        return pos_ < text_.size() ? text_[pos_] : '\0';
    }

    bool accept(char const c)
    {
        if (peek() != c) return false;
        ++pos_;
        return true;
    }

    void expect(char const c)
    {
        if (!accept(c)) fail(std::string("expected '") + c + "'");
    }

    static bool word_char(char const c)
    {
        return std::isalnum(static_cast<unsigned char>(c)) || c == '-' || c == '_' || c == '.';
    }

    // A keyword: matches only a whole word, so `table` does not swallow the head of `table1`.
    bool accept_word(char const* word)
    {
        skip();
        std::size_t n = 0;
        while (word[n]) ++n;
        if (text_.compare(pos_, n, word) != 0) return false;
        if (pos_ + n < text_.size() && word_char(text_[pos_ + n])) return false;
        pos_ += n;
        return true;
    }

    void expect_word(char const* word)
    {
        if (!accept_word(word)) fail(std::string("expected '") + word + "'");
    }

    // A name: the reference's usable_string, +(alnum | '-' | '_') (bif.hpp:196-197), plus '.' and
    // the quoted form "..." that JavaBayes and MSR DSC files use.
    std::string name()
    {
        skip();
        std::string out;
        if (pos_ < text_.size() && text_[pos_] == '"') {
            std::size_t const close = text_.find('"', pos_ + 1);
            if (close == std::string::npos) fail("unterminated string");
            out = text_.substr(pos_ + 1, close - pos_ - 1);
            pos_ = close + 1;
            return out;
        }
        while (pos_ < text_.size() && word_char(text_[pos_])) out.push_back(text_[pos_++]);
        if (out.empty()) fail("expected a name");
        return out;
    }

    bool at_number()
    {
        char const c = peek();
        return std::isdigit(static_cast<unsigned char>(c)) || c == '.' || c == '-' || c == '+';
    }

    double number()
    {
        skip();
        char const* const begin = text_.c_str() + pos_;
        char* end = nullptr;
        double const value = std::strtod(begin, &end);
        if (end == begin) fail("expected a number");
        pos_ += static_cast<std::size_t>(end - begin);
        return value;
    }

    std::size_t unsigned_integer()
    {
        skip();
        std::size_t value = 0, digits = 0;
        while (pos_ < text_.size() && std::isdigit(static_cast<unsigned char>(text_[pos_]))) {
            value = value * 10 + static_cast<std::size_t>(text_[pos_++] - '0');
            ++digits;
        }
        if (!digits) fail("expected an unsigned integer");
        return value;
    }

    // Consume up to and including the next `stop` character (property lines, unknown attributes).
    void skip_past(char const stop)
    {
        std::size_t const at = text_.find(stop, pos_);
        if (at == std::string::npos) fail(std::string("expected '") + stop + "'");
        pos_ = at + 1;
    }

    [[noreturn]] void fail(std::string const& what) const
    {
        std::size_t line = 1;
        for (std::size_t i = 0; i < pos_ && i < text_.size(); ++i)
            if (text_[i] == '\n') ++line;
        throw std::runtime_error(std::string("cannot parse ") + format_ + ": " + what + " at line " + std::to_string(line));
    }

private:
    std::string text_;
    char const* format_;
    std::size_t pos_ = 0;
};

} // namespace serializer
} // namespace bn

#endif // BNB200_BAYESIAN_SERIALIZER_TEXT_SCANNER_HPP
