"""unicore.data.Dictionary restated (used at reference models/mm_model.py:435-438,
data/conformer.py:65-66): one symbol per line, indices in file order."""


class Dictionary:
    def __init__(self, *, bos="[CLS]", pad="[PAD]", eos="[SEP]", unk="[UNK]"):
        self.bos_word, self.pad_word, self.eos_word, self.unk_word = bos, pad, eos, unk
        self.symbols, self.indices, self.specials = [], {}, set()
        for w in (bos, pad, eos, unk):
            self.specials.add(w)

    def __len__(self):
        return len(self.symbols)

    def __getitem__(self, i):
        return self.symbols[i] if i < len(self.symbols) else self.unk_word

    def index(self, sym):
        return self.indices.get(sym, self.indices.get(self.unk_word))

    def add_symbol(self, word, is_special=False):
        if is_special:
            self.specials.add(word)
        if word not in self.indices:
            self.indices[word] = len(self.symbols)
            self.symbols.append(word)
        return self.indices[word]

    def bos(self):
        return self.index(self.bos_word)

    def pad(self):
        return self.index(self.pad_word)

    def eos(self):
        return self.index(self.eos_word)

    def unk(self):
        return self.index(self.unk_word)

    @classmethod
    def load(cls, path):
        d = cls()
        with open(path, "r", encoding="utf-8") as fh:
            for line in fh:
                tok = line.strip().split()
                if tok:
                    d.add_symbol(tok[0])
        return d
