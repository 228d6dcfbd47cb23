from insilicoseq_b200.app import main

if __name__ == "__main__":
    main()
