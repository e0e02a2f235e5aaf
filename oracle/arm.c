/* placeholder; arm model restatement lands with the arm collision kernel */
